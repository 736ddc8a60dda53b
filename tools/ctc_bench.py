"""CTC loss+grad kernel timing (us per batch) at the BASELINE shapes; HTRVT_CTC_SLOW=1 times the log-space path."""
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
for (B, T, C, lo, hi) in [(128, 128, 80, 16, 64), (1024, 128, 80, 16, 64), (4096, 128, 80, 16, 64), (128, 256, 90, 64, 200)]:
    rs = np.random.RandomState(0)
    x = torch.randn(B, T, C, device="cuda")
    tl = torch.from_numpy(rs.randint(lo, hi + 1, size=B).astype(np.int32))
    tg = torch.from_numpy(rs.randint(1, C, size=int(tl.sum())).astype(np.int32)).cuda()
    tld = tl.cuda()
    mtl = int(tl.max())
    f = lambda: ops.ctc_loss_grad(x, tg, None, tld, layout="btc", is_logprob=False, max_target_len=mtl)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    n0 = lib.htrvt_ctc_fallback_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    byts = 2.0 * B * T * C * 4 + float(tl.sum()) * 4 + 12 * B
    if os.environ.get("HTRVT_CTC_DEBUG"):
        import ctypes
        buf = (ctypes.c_longlong * 16)()
        lib.htrvt_ctc_debug_stamps.argtypes = [ctypes.c_void_p]
        lib.htrvt_ctc_debug_stamps(ctypes.cast(buf, ctypes.c_void_p))
        st = list(buf)
        print("  clocks from CTA0 start: sync0a %d, phase0 %d, alpha %d, beta %d, phase2(w5) %d, end %d  (L0=%d)" %
              (st[1] - st[0], st[2] - st[0], st[3] - st[0], st[4] - st[0], st[5] - st[0], st[6] - st[0], int(tl[0])))
        torch.cuda.synchronize()
    print("B=%d T=%d C=%d L<=%d: %.1f us/batch  %.0f GB/s algorithmic  fallbacks %d" %
          (B, T, C, hi, us, byts / us / 1e3, lib.htrvt_ctc_fallback_count() - n0))
