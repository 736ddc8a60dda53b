"""How long does the host need to ENQUEUE one training step (python + ctypes + launches) vs the GPU time?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from importlib import import_module
import bench
import htrvt_b200 as h
H = import_module("htr-vt_b200.model.HTR_VT")
dev = torch.device("cuda", 0)
B = 128
model = H.create_model(bench.NB_CLS, [bench.IMG_H, bench.IMG_W]).to(dev).train()
crit = h.CTCLoss(reduction="none", zero_infinity=True)
img, tg, tl = [t.to(dev) for t in bench.synth_batch(B, 0)]
def step():
    for p in model.parameters(): p.grad = None
    preds = model(img, bench.MASK_RATIO, bench.MAX_SPAN, use_masking=True).float()
    ps = torch.full((B,), preds.size(1), dtype=torch.int32, device=dev)
    loss = crit(preds.permute(1, 0, 2).log_softmax(2), tg, ps, tl).mean()
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("enqueue %.2f ms, total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
