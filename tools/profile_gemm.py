"""Isolated launches of representative tap-GEMM problems (for ncu source-level analysis and quick timing)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import htrvt_b200  # noqa: F401,E402

o = import_module("htr-vt_b200.ops")
dev = "cuda"
torch.manual_seed(0)
M = 16384
x768 = torch.randn(M, 768, device=dev).bfloat16()
x3072 = torch.randn(M, 3072, device=dev).bfloat16()
w_proj = torch.randn(768, 768, device=dev).bfloat16()
w_fc1 = torch.randn(3072, 768, device=dev).bfloat16()
w_fc2 = torch.randn(768, 3072, device=dev).bfloat16()
w_qkv = torch.randn(2304, 768, device=dev).bfloat16()
b768, b3072, b2304 = torch.randn(768, device=dev), torch.randn(3072, device=dev), torch.randn(2304, device=dev)
res = torch.randn(M, 768, device=dev)
y32 = torch.empty(M, 768, device=dev)
a_out, u_out = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16), torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
qkv = torch.empty(3, 128, 6, 128, 128, device=dev, dtype=torch.bfloat16)
xl1 = torch.randn(128, 8, 512, 192, device=dev).bfloat16()
wl1 = torch.randn(192, 9, 192, device=dev).bfloat16()
stats = torch.zeros(4 * 148, 2, 192, device=dev)
stats3 = torch.zeros(4 * 148, 2, 768, device=dev)
yl1 = torch.empty(128, 8, 512, 192, device=dev, dtype=torch.bfloat16)
gl1 = torch.zeros(192, 192, 3, 3, device=dev)
xl3 = torch.randn(128, 2, 128, 768, device=dev).bfloat16()
wl3 = torch.randn(768, 9, 768, device=dev).bfloat16()
yl3 = torch.empty(128, 2, 128, 768, device=dev, dtype=torch.bfloat16)
gl3 = torch.zeros(768, 768, 3, 3, device=dev)
g_fc1 = torch.zeros(3072, 768, device=dev)
xl2 = torch.randn(128, 4, 256, 384, device=dev).bfloat16()
yl2 = torch.randn(128, 4, 256, 384, device=dev).bfloat16()
gl2 = torch.zeros(384, 384, 3, 3, device=dev)

wl1T = wl1.permute(2, 1, 0).contiguous()
wl2 = torch.randn(384, 9, 384, device=dev).bfloat16()
wl2T = wl2.permute(2, 1, 0).contiguous()
yl2b = torch.empty(128, 4, 256, 384, device=dev, dtype=torch.bfloat16)
gt1 = torch.zeros(192, 9, 192, device=dev)
gt2 = torch.zeros(384, 9, 384, device=dev)
gt3 = torch.zeros(768, 9, 768, device=dev)

# operand major-ness experiment: the same 3072 x 768 x 16384 contraction with (A,B) = (K,K), (K,MN), (MN,MN)
aK = torch.randn(3072, 16384, device=dev).bfloat16()
bK = torch.randn(768, 16384, device=dev).bfloat16()
bMN = torch.randn(16384, 768, device=dev).bfloat16()
oKK = torch.empty(3072, 768, device=dev)
cases = {
    "mm_KK": (lambda: o.gemm_tn(aK, bK, oKK), 2.0 * 3072 * 768 * 16384),
    "mm_KMN": (lambda: o.gemm_nn(aK, bMN, oKK), 2.0 * 3072 * 768 * 16384),
    "mm_MNMN": (lambda: o.linear_wgrad(x3072, x768, g_fc1), 2.0 * 3072 * 768 * 16384),
    "proj": (lambda: o.gemm_tn(x768, w_proj, a_out.view(-1)[: M * 768].view(M, 768), bias=b768), 2.0 * M * 768 * 768),
    "fc1": (lambda: o.gemm_tn(x768, w_fc1, a_out, bias=b3072), 2.0 * M * 3072 * 768),
    "fc2": (lambda: o.gemm_tn(x3072, w_fc2, u_out.view(-1)[: M * 768].view(M, 768), bias=b768), 2.0 * M * 3072 * 768),
    "qkv": (lambda: o.gemm_tn(x768, w_qkv, a_out.view(-1)[: M * 2304].view(M, 2304), bias=b2304), 2.0 * M * 2304 * 768),
    "fc1_dgrad": (lambda: o.gemm_nn(x3072, w_fc1, yl1.view(-1)[: M * 768].view(M, 768)), 2.0 * M * 3072 * 768),
    "fc1_wgrad": (lambda: o.linear_wgrad(x3072, x768, g_fc1), 2.0 * M * 3072 * 768),
    "l1_fwd_stats": (lambda: o.conv_fwd(xl1, wl1, 3, 1, 1, y=yl1, stats=stats), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l1_fwd_nostats": (lambda: o.conv_fwd(xl1, wl1, 3, 1, 1, y=yl1), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l1_fwd_nostore": (lambda: o.conv_fwd(xl1, wl1, 3, 1, 1, y=yl1, nostore=True), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l1_dgrad": (lambda: o.conv_dgrad(yl1, wl1, (128, 8, 512, 192), 3, 1, 1, dx=xl1), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l1_dgrad_acc": (lambda: o.conv_dgrad(yl1, wl1, (128, 8, 512, 192), 3, 1, 1, dx=xl1, accumulate=True), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l1_dgrad_T": (lambda: o.conv_dgrad(yl1, wl1, (128, 8, 512, 192), 3, 1, 1, dx=xl1, w_t=wl1T), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l2_dgrad": (lambda: o.conv_dgrad(yl2, wl2, (128, 4, 256, 384), 3, 1, 1, dx=xl2), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l2_dgrad_T": (lambda: o.conv_dgrad(yl2, wl2, (128, 4, 256, 384), 3, 1, 1, dx=xl2, w_t=wl2T), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l2_fwd": (lambda: o.conv_fwd(xl2, wl2, 3, 1, 1, y=yl2b), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l1_wgrad_acc": (lambda: o.conv_wgrad_acc(yl1, xl1, 3, 1, 1, gt1), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l2_wgrad_acc": (lambda: o.conv_wgrad_acc(yl2, xl2, 3, 1, 1, gt2), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l3_wgrad_acc": (lambda: o.conv_wgrad_acc(yl3, xl3, 3, 1, 1, gt3), 2.0 * 128 * 2 * 128 * 768 * 6912),
    "l1_wgrad_t": (lambda: o.conv_wgrad_acc_t(yl1, xl1, 3, 1, 1, gt1.view(9, 192, 192)), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l2_wgrad_t": (lambda: o.conv_wgrad_acc_t(yl2, xl2, 3, 1, 1, gt2.view(9, 384, 384)), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l3_wgrad_t": (lambda: o.conv_wgrad_acc_t(yl3, xl3, 3, 1, 1, gt3.view(9, 768, 768)), 2.0 * 128 * 2 * 128 * 768 * 6912),
    "l1_wgrad_w": (lambda: o.conv_wgrad_acc_w(yl1, xl1, 1, gt1.view(3, 3, 3, 64, 192)), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l2_wgrad_w": (lambda: o.conv_wgrad_acc_w(yl2, xl2, 1, gt2.view(3, 6, 3, 64, 384)), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l3_wgrad_w": (lambda: o.conv_wgrad_acc_w(yl3, xl3, 1, gt3.view(3, 12, 3, 64, 768)), 2.0 * 128 * 2 * 128 * 768 * 6912),
    "l1_wgrad": (lambda: o.conv_wgrad(yl1, xl1, 3, 1, 1, gl1), 2.0 * 128 * 8 * 512 * 192 * 1728),
    "l3_fwd_stats": (lambda: o.conv_fwd(xl3, wl3, 3, 1, 1, y=yl3, stats=stats3), 2.0 * 128 * 2 * 128 * 768 * 6912),
    "l2_wgrad": (lambda: o.conv_wgrad(yl2, xl2, 3, 1, 1, gl2), 2.0 * 128 * 4 * 256 * 384 * 3456),
    "l3_wgrad": (lambda: o.conv_wgrad(yl3, xl3, 3, 1, 1, gl3), 2.0 * 128 * 2 * 128 * 768 * 6912),
}
sel = sys.argv[1].split(",") if len(sys.argv) > 1 else list(cases)
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for name in sel:
    fn, fl = cases[name]
    for _ in range(2 if n_iter > 1 else 1):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = n_iter
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-14s %8.1f us  %7.1f TFLOP/s" % (name, ms * 1e3, fl / ms / 1e9))
