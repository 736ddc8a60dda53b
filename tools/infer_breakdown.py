"""Per-op breakdown of the eval-mode forward + greedy decode at 512 lines (BASELINE config 4 share)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from importlib import import_module
import bench
import htrvt_b200 as h
ops = import_module("htr-vt_b200.ops")
H = import_module("htr-vt_b200.model.HTR_VT")
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(123)
model = H.create_model(bench.NB_CLS, [bench.IMG_H, bench.IMG_W]).to(dev).eval()
conv = h.CTCLabelConverter("".join(chr(33 + i) for i in range(bench.NB_CLS - 1)))
img = bench.synth_batch(B, 1)[0].to(dev)
def once():
    with torch.no_grad():
        return conv.decode_logits(model(img).float())
for _ in range(3): once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): once()
e1.record(); torch.cuda.synchronize()
print("eval forward + decode: %.3f ms per %d lines = %.0f img/s" % (e0.elapsed_time(e1) / 5, B, B / (e0.elapsed_time(e1) / 5) * 1e3))
ops.PROFILE = []
once()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
by = {}
for name, fl, a, b in prof:
    d = by.setdefault(name, [0.0, 0.0, 0]); d[0] += a.elapsed_time(b); d[1] += fl; d[2] += 1
tot = sum(v[0] for v in by.values())
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print("%-20s %3d launches %8.3f ms %5.1f%%  %s" % (k, v[2], v[0], 100 * v[0] / tot, ("%.0f TFLOP/s" % (v[1] / v[0] / 1e9)) if v[1] else ""))
print("sum of ops %.3f ms" % tot)
