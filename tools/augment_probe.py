"""Timing probe: SameTrCollate's pixel work on the GPU (csrc/augment.cu), 128 lines of 64 x 512, all three stages on."""
import os
import sys
import time
import types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from importlib import import_module
aug = import_module("htr-vt_b200.augment")
ops = import_module("htr-vt_b200.ops")

args = types.SimpleNamespace(proj=8.0, dila_ero_max_kernel=3, dila_ero_iter=1, jitter_contrast=0.4,
                             jitter_brightness=0.4, jitter_saturation=0.4, jitter_hue=0.2)
for B, H, W in ((128, 64, 512), (128, 64, 1024)):
    x = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device="cuda")
    seed = 0
    while True:
        seed += 1
        np.random.seed(seed); torch.manual_seed(seed)
        p = aug.draw_collate_params(B, H, W, args)
        if all(p[k] is not None for k in ("warp", "morph", "jitter")):
            break
    t0 = time.time()
    for _ in range(20):
        np.random.seed(seed); torch.manual_seed(seed)
        aug.draw_collate_params(B, H, W, args)
    t_draw = (time.time() - t0) / 20
    t0 = time.time()
    for _ in range(20):
        rec, morph = aug.pack_params(p, B, H, W)
    t_pack = (time.time() - t0) / 20
    recd = torch.from_numpy(rec).cuda()
    for _ in range(3):
        ops.augment_lines(x, recd, morph)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(20):
        ops.augment_lines(x, recd, morph)
    torch.cuda.synchronize()
    t_k = (time.time() - t0) / 20
    print("B %d %dx%d: kernel %.1f us, host draws %.2f ms, record packing %.2f ms per batch" %
          (B, H, W, t_k * 1e6, t_draw * 1e3, t_pack * 1e3))
