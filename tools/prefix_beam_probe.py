"""Timing probe: CTC prefix beam search (csrc/prefix_beam.cu) vs the K-best path kernel on 512 lines of T=128, C=80."""
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from importlib import import_module
ops = import_module("htr-vt_b200.ops")


def t_us(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.time() - t) / n * 1e6


for T, B, C in ((128, 512, 80), (256, 512, 90)):
    lp = torch.randn(T, B, C, device="cuda").log_softmax(2)
    for K in (1, 5, 8, 16):
        print("T %d C %d prefix K %2d: %8.1f us" % (T, C, K, t_us(lambda: ops.ctc_prefix_beam(lp, K))))
    print("T %d C %d kbest  K  5: %8.1f us" % (T, C, t_us(lambda: ops.ctc_kbest_paths(lp, 5))))
