"""Lane-group CTC kernel: parity against float64 torch CTC, flagged-sequence counts, and us/batch of the two kernels
(mode 0 CTA per sequence, 1 lane group per sequence) over batch sizes.  CTC_NCU=1: one B = 4096 call only (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from importlib import import_module  # noqa: E402
import htrvt_b200  # noqa: F401,E402

ops = import_module("htr-vt_b200.ops")
lib = import_module("htr-vt_b200._lib").lib()


def ref64(logits, tg, il, tl):
    lg = torch.from_numpy(logits).double().requires_grad_(True)
    lp = lg.permute(1, 0, 2).log_softmax(2)
    nll = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg).long(), torch.from_numpy(il).long(),
                                       torch.from_numpy(tl).long(), blank=0, reduction="none", zero_infinity=True)
    nll.sum().backward()
    return nll.detach().numpy(), lg.grad.numpy()


def make(B, T, C, lo, hi, scale=1.0, seed=0):
    rs = np.random.RandomState(seed)
    x = (rs.randn(B, T, C) * scale).astype(np.float32)
    tl = rs.randint(lo, hi + 1, size=B).astype(np.int32)
    tg = rs.randint(1, C, size=int(tl.sum())).astype(np.int32)
    return x, tg, tl


def run(x, tg, tl, mode, il=None):
    prev = lib.htrvt_ctc_set_mode(mode)
    xd = torch.from_numpy(x).cuda() if isinstance(x, np.ndarray) else x
    nll, g = ops.ctc_loss_grad(xd, torch.from_numpy(tg).cuda(), None if il is None else torch.from_numpy(il).cuda(),
                               torch.from_numpy(tl).cuda(), layout="btc", is_logprob=False, max_target_len=int(tl.max()))
    torch.cuda.synchronize()
    lib.htrvt_ctc_set_mode(prev)
    return nll.cpu().numpy(), g.cpu().numpy()


def timeit(x, tg, tl, mode, n=30):
    prev = lib.htrvt_ctc_set_mode(mode)
    xd, tgd, tld, mtl = torch.from_numpy(x).cuda(), torch.from_numpy(tg).cuda(), torch.from_numpy(tl).cuda(), int(tl.max())
    f = lambda: ops.ctc_loss_grad(xd, tgd, None, tld, layout="btc", is_logprob=False, max_target_len=mtl)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    lib.htrvt_ctc_set_mode(prev)
    return e0.elapsed_time(e1) / n * 1e3


if os.environ.get("CTC_NCU"):
    x, tg, tl = make(int(os.environ.get("CTC_B", "4096")), 128, 80, 16, 64)
    for _ in range(3):
        run(x, tg, tl, 1)
    print("done")
    sys.exit(0)

for (B, T, C, lo, hi, scale) in [(64, 128, 80, 16, 64, 1.0), (64, 128, 80, 1, 32, 4.0), (32, 256, 90, 64, 200, 1.0),
                                 (16, 128, 228, 1, 64, 8.0), (40, 64, 37, 0, 20, 2.0)]:
    x, tg, tl = make(B, T, C, lo, hi, scale, seed=B + T)
    il = np.full(B, T, dtype=np.int32)
    f0 = lib.htrvt_ctc_flagged_count()
    nll, g = run(x, tg, tl, 1)
    f1 = lib.htrvt_ctc_flagged_count()
    rn, rg = ref64(x, tg, il, tl)
    en = np.abs(nll - rn) / np.maximum(np.abs(rn), 1e-9)
    eg = np.abs(g - rg)
    tolg = eg - (1e-5 + 1e-4 * np.abs(rg))
    print("parity B=%d T=%d C=%d L=%d..%d scale %.0f: nll relerr %.2e, grad abs err %.2e (allclose margin %.2e), row sums %.2e, "
          "flagged %d" % (B, T, C, lo, hi, scale, en.max(), eg.max(), tolg.max(), np.abs(g.sum(2)).max(), f1 - f0))

for B in (128, 256, 512, 1024, 2048, 4096, 8192):
    x, tg, tl = make(B, 128, 80, 16, 64)
    byts = 2.0 * B * 128 * 80 * 4 + float(tl.sum()) * 4 + 12 * B
    f0 = lib.htrvt_ctc_flagged_count()
    us = [timeit(x, tg, tl, m) for m in (0, 1)]
    print("B=%5d T=128 C=80: CTA %.1f us, group %.1f us (%.0f GB/s algorithmic, flagged %d)"
          % (B, us[0], us[1], byts / us[1] / 1e3, lib.htrvt_ctc_flagged_count() - f0))
x, tg, tl = make(128, 256, 90, 64, 200)
print("B=128 T=256 C=90 L<=200: CTA %.1f us, group %.1f us" % (timeit(x, tg, tl, 0), timeit(x, tg, tl, 1)))
x, tg, tl = make(2048, 256, 90, 64, 200)
print("B=2048 T=256 C=90 L<=200: CTA %.1f us, group %.1f us" % (timeit(x, tg, tl, 0), timeit(x, tg, tl, 1)))
