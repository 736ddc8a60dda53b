"""One call of the CTC loss+grad entry point at B = 4096 (warp-per-sequence kernel) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import htrvt_b200  # noqa: F401,E402

ops = import_module("htr-vt_b200.ops")
B, T, C = int(os.environ.get("CTC_B", "4096")), 128, 80
rs = np.random.RandomState(0)
x = torch.randn(B, T, C, device="cuda")
tl = torch.from_numpy(rs.randint(16, 65, size=B).astype(np.int32))
tg = torch.from_numpy(rs.randint(1, C, size=int(tl.sum())).astype(np.int32)).cuda()
tld = tl.cuda()
for _ in range(3):
    ops.ctc_loss_grad(x, tg, None, tld, layout="btc", is_logprob=False, max_target_len=int(tl.max()))
torch.cuda.synchronize()
print("done")
