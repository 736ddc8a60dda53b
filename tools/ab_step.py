"""A/B timing of the training step inside ONE process (same box, same clocks): alternates blocks of steps with a module
flag off / on and prints the per-variant medians.  Usage: python tools/ab_step.py engine.FUSE_BN_BWD [rounds] [steps]
(a callable attribute, e.g. ops.set_pdl, is called with False / True instead of being assigned)"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import bench  # noqa: E402

modname, attr = sys.argv[1].rsplit(".", 1)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch_, dist, world, rank, local, dev, h, ops, H = bench.setup_ours()
mod = import_module("htr-vt_b200." + modname)
B = 128
torch.manual_seed(123)
model = H.create_model(bench.NB_CLS, [bench.IMG_H, bench.IMG_W]).to(dev).train()
criterion = h.CTCLoss(reduction="none", zero_infinity=True).to(dev)
img, tg, tl = [t.to(dev) for t in bench.synth_batch(B, seed=0)]
params = [p for p in model.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    preds = model(img, bench.MASK_RATIO, bench.MAX_SPAN, use_masking=True).float()
    ps = torch.full((B,), preds.size(1), dtype=torch.int32, device=dev)
    loss = criterion(preds.permute(1, 0, 2).log_softmax(2), tg, ps, tl).mean()
    loss.backward()


_target = getattr(mod, attr)
if callable(_target):
    def _set(v):
        _target(v)
else:
    def _set(v):
        setattr(mod, attr, v)
res = {False: [], True: []}
for v in (False, True):
    _set(v)
    for _ in range(3):
        step()
for r in range(rounds):
    for v in (False, True) if r % 2 == 0 else (True, False):
        _set(v)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[v].append(e0.elapsed_time(e1) / steps)
for v in (False, True):
    print("%s=%s: median %.3f ms/step  (%s)" % (sys.argv[1], v, statistics.median(res[v]), " ".join("%.2f" % x for x in res[v])))
