"""Where the end-to-end step loses time against the device-resident step: H2D copy, loss.item() bubble, enqueue rate."""
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
torch.manual_seed(123)
model = H.create_model(bench.NB_CLS, [bench.IMG_H, bench.IMG_W]).to(dev).train()
crit = h.CTCLoss(reduction="none", zero_infinity=True)
img_h, tg_h, tl_h = [t.pin_memory() for t in bench.synth_batch(B, 0)]
img_d, tg_d, tl_d = img_h.to(dev), tg_h.to(dev), tl_h.to(dev)
params = [p for p in model.parameters() if p.requires_grad]
ps = torch.full((B,), bench.IMG_W // 4, dtype=torch.int32, device=dev)
side = torch.cuda.Stream()
pin_loss = torch.zeros(64, pin_memory=True)

def fb(image, text, length):
    for p in params:
        p.grad = None
    preds = model(image, bench.MASK_RATIO, bench.MAX_SPAN, use_masking=True).float()
    loss = crit(preds.permute(1, 0, 2).log_softmax(2), text, ps, length).mean()
    loss.backward()
    return loss

state = {"i": 0, "next": None}
def v_resident(): fb(img_d, tg_d, tl_d)
def v_resident_item(): return fb(img_d, tg_d, tl_d).item()
def v_h2d(): fb(img_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True), tl_h.to(dev, non_blocking=True))
def v_h2d_item(): return fb(img_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True), tl_h.to(dev, non_blocking=True)).item()
def v_h2d_async_read():
    loss = fb(img_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True), tl_h.to(dev, non_blocking=True))
    pin_loss[state["i"] % 64].copy_(loss.detach(), non_blocking=True)
    state["i"] += 1
def v_prefetch_async_read():
    cur = state["next"]
    if cur is None:
        cur = (img_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True), tl_h.to(dev, non_blocking=True))
    with torch.cuda.stream(side):
        state["next"] = (img_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True), tl_h.to(dev, non_blocking=True))
    loss = fb(*cur)
    torch.cuda.current_stream().wait_stream(side)
    pin_loss[state["i"] % 64].copy_(loss.detach(), non_blocking=True)
    state["i"] += 1

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    t_enq = (time.perf_counter() - t0) / n * 1e3
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_enq

for name, fn in [("resident", v_resident), ("resident+item", v_resident_item), ("h2d", v_h2d), ("h2d+item", v_h2d_item),
                 ("h2d+async_read", v_h2d_async_read), ("prefetch+async_read", v_prefetch_async_read)]:
    ms, enq = timed(fn)
    print("%-22s %.3f ms/step   host enqueue %.3f ms/step" % (name, ms, enq))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): img_h.to(dev, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D of one image batch (%.1f MB): %.3f ms" % (img_h.numel() * 4 / 1e6, e0.elapsed_time(e1) / 10))
