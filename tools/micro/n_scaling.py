"""Does tcgen05.mma throughput scale linearly with the tile's N?  The layer-1 3x3 convolution (Cin = 192, 128 x 8 x 512
pixels: the A operand is re-read 9x from L2, so the kernel is not HBM-bound) with Cout = 128 / 192 / 256 / 384 / 512."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from importlib import import_module
import htrvt_b200  # noqa
o = import_module("htr-vt_b200.ops")
x = torch.randn(128, 8, 512, 192, device="cuda").bfloat16()
for N in (128, 192, 256, 384, 512):
    w = torch.randn(N, 9, 192, device="cuda").bfloat16()
    y = torch.empty(128, 8, 512, N, device="cuda", dtype=torch.bfloat16)
    f = lambda: o.conv_fwd(x, w, 3, 1, 1, y=y)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("Cout=%4d  %8.1f us  %7.1f TFLOP/s" % (N, ms * 1e3, 2.0 * 128 * 8 * 512 * N * 1728 / ms / 1e9))
