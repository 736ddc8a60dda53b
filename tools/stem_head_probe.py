"""Stem-head forward kernels (tensor pipe / FP32 pipe) at the training and inference shapes: us per call; for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import htrvt_b200  # noqa: F401,E402

ops = import_module("htr-vt_b200.ops")
lib = import_module("htr-vt_b200._lib").lib()
C, H, W = 192, 64, 512
w = torch.randn(C, 1, 3, 3, device="cuda") * 0.4
st = (None, None, torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.3)
for B, code in ((128, True), (512, False)):
    x = torch.randn(B, H, W, device="cuda")
    for mode in (0, 1):
        prev = lib.htrvt_stem_head_set_mode(mode)
        f = lambda: ops.stem_head_fwd(x, w, st, code, out_dtype=torch.float16, want_bf16=code)
        for _ in range(2):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 1 if os.environ.get("PROBE_NCU") else 10
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        lib.htrvt_stem_head_set_mode(prev)
        print("B=%d code=%s %s: %.1f us" % (B, code, "tensor pipe" if mode else "fp32 pipe", e0.elapsed_time(e1) / n * 1e3))
