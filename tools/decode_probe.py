#!/usr/bin/env python
"""Kernel time of the greedy decode (512 x 128 x 80 fp32 logits): 64 calls in one CUDA graph over 8 rotating buffers
(168 MB > L2), replayed 20 times.  Same-box A/B of two builds: copy each .so over htr-vt_b200/libhtrvt_b200.so and run
this in a fresh process (tools/ab_so.sh pattern)."""
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ops = import_module("htr-vt_b200.ops")
dev = torch.device("cuda:0")
B, T, C = 512, 128, 80
bufs = [torch.randn(B, T, C, device=dev) for _ in range(8)]
ops.greedy_decode_ids(bufs[0], C)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for i in range(64):
            ops.greedy_decode_ids(bufs[i % 8], C)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(20):
        g.replay()
    e1.record(side)
    side.synchronize()
us = e0.elapsed_time(e1) / (20 * 64) * 1e3
by = B * T * C * 4.0 + B * T * 4.0 + B * 4.0
print("%s greedy_decode %.2f us  %.0f GB/s" % (sys.argv[1] if len(sys.argv) > 1 else "", us, by / us / 1e3))
