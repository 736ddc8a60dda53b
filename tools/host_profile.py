"""cProfile of the host side of the training step (where the 11 ms of enqueue time per step go)."""
import cProfile, pstats, os, sys
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
img_d, tg_d, tl_d = [t.to(dev) for t in bench.synth_batch(B, 0)]
params = [p for p in model.parameters() if p.requires_grad]
ps = torch.full((B,), bench.IMG_W // 4, dtype=torch.int32, device=dev)
def fb():
    for p in params:
        p.grad = None
    preds = model(img_d, bench.MASK_RATIO, bench.MAX_SPAN, use_masking=True).float()
    loss = crit(preds.permute(1, 0, 2).log_softmax(2), tg_d, ps, tl_d).mean()
    loss.backward()
for _ in range(3): fb()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10): fb()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(35)
