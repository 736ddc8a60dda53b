#!/usr/bin/env python
"""The optimizer launches of one SAM iteration alone (first_step, second_step with the fused AdamW, EMA update):
    python tools/optim_probe.py            device time with the host enqueued ahead (bench.py's optimizer_ms)
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/optim_launches.csv python tools/optim_probe.py ncu      launch list of ONE optimizer pass
"""
import os
import sys
from importlib import import_module

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H = import_module("htr-vt_b200.model.HTR_VT")
SAMm = import_module("htr-vt_b200.utils.sam")
Um = import_module("htr-vt_b200.utils.utils")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = H.create_model(80, [64, 512]).to(dev).train()
opt = SAMm.SAM(model.parameters(), torch.optim.AdamW, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.5)
ema = Um.ModelEma(model, 0.9999)
for p in model.parameters():
    p.grad = torch.randn_like(p) * 1e-3


def opt_only():
    opt.first_step(zero_grad=False)
    opt.second_step(zero_grad=False)
    ema.update(model, num_updates=10)


opt_only()
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "ncu":
    torch.cuda.cudart().cudaProfilerStart()
    opt_only()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
big = torch.empty(1 << 28, device=dev)
ts = []
for _ in range(7):
    for _ in range(40):                     # ~20 ms of queued fills: the host finishes enqueueing long before the device
        big.fill_(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt_only()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
import time
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    opt_only()
host_ms = (time.perf_counter() - t0) / 20 * 1e3        # enqueue cost of one optimizer pass (the stream never drains)
torch.cuda.synchronize()
n = sum(p.numel() for p in model.parameters())
by = n * (4 + 16 + 28 + 12)
print("%s optimizer passes %.3f ms (median of 7, min %.3f)  %.0f GB/s algorithmic; host enqueue %.3f ms per pass" % (
    sys.argv[1] if len(sys.argv) > 1 else "", ts[3], ts[0], by / ts[3] / 1e6, host_ms))
