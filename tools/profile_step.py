"""One training step of the bench workload inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import bench  # noqa: E402
import htrvt_b200 as h  # noqa: E402

H = import_module("htr-vt_b200.model.HTR_VT")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(123)
model = H.create_model(bench.NB_CLS, [bench.IMG_H, bench.IMG_W]).to(dev).train()
crit = h.CTCLoss(reduction="none", zero_infinity=True)
img, tg, tl = [t.to(dev) for t in bench.synth_batch(B, 0)]


def step():
    for p in model.parameters():
        p.grad = None
    preds = model(img, bench.MASK_RATIO, bench.MAX_SPAN, use_masking=True).float()
    ps = torch.full((B,), preds.size(1), dtype=torch.int32, device=dev)
    loss = crit(preds.permute(1, 0, 2).log_softmax(2), tg, ps, tl).mean()
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
conv = h.CTCLabelConverter("".join(chr(33 + i) for i in range(bench.NB_CLS - 1)))
logits = torch.randn(512, 128, bench.NB_CLS, device=dev)
torch.cuda.profiler.start()
step()
conv.decode_logits(logits)            # greedy argmax + collapse kernel (inference path), 512 lines
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, B =", B)
