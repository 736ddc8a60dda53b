"""GPU parity of the augmentation kernel (csrc/augment.cu) behind the reference's SameTrCollate signature:
committed goldens from the reference's own collate, and the numpy oracle on seeded random batches (all gates, morph
shapes up to 3 x 3 x 2 iterations, the anti-aliased tall-warp case, wide lines, odd sizes).  Bit exact: uint8 work."""
import os
import sys
import types
from importlib import import_module

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import augment_oracle as A  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(ROOT, "tests", "golden")


def _args(v=(8.0, 3, 1, 0.4, 0.4, 0.4, 0.2)):
    return types.SimpleNamespace(proj=float(v[0]), dila_ero_max_kernel=int(v[1]), dila_ero_iter=int(v[2]),
                                 jitter_brightness=float(v[3]), jitter_contrast=float(v[4]),
                                 jitter_saturation=float(v[5]), jitter_hue=float(v[6]))


def _lines(rs, B, H, W):
    x = np.full((B, H, W), 255, dtype=np.uint8)
    x -= rs.randint(0, 16, (B, H, W)).astype(np.uint8)
    for b in range(B):
        for _ in range(30):
            r0, c0 = rs.randint(0, H), rs.randint(0, W)
            hh, ww = rs.randint(1, max(2, H // 3)), rs.randint(1, 9)
            x[b, r0:r0 + hh, c0:c0 + ww] = rs.randint(0, 140)
    return x


def test_collate_matches_reference_goldens():
    aug = import_module("htr-vt_b200.augment")
    g = np.load(os.path.join(G, "augment_cases.npz"))
    args = _args(g["args"])
    imgs = g["images"]
    batch = [(imgs[i], "l%d" % i) for i in range(len(imgs))]
    for seed, want in zip(g["seeds"], g["outputs"]):
        np.random.seed(int(seed)); torch.manual_seed(int(seed))
        out, labels = aug.SameTrCollate(batch, args, as_uint8=True)
        assert out.dtype == torch.uint8 and out.shape == (len(imgs), 1, 64, imgs.shape[-1]) and out.is_cuda
        assert list(labels) == ["l%d" % i for i in range(len(imgs))]
        got = out[:, 0].cpu().numpy()
        assert np.array_equal(got, want), (int(seed), int((got != want).sum()))
        np.random.seed(int(seed)); torch.manual_seed(int(seed))
        outf, _ = aug.SameTrCollate(batch, args)                       # the reference's return type: float in [0, 1]
        assert outf.dtype == torch.float32
        nbad = int((outf[:, 0].cpu().numpy() != want.astype(np.float32) / np.float32(255.)).sum())
        assert nbad == 0, (int(seed), nbad)


@pytest.mark.parametrize("H,W,B,it", [(64, 512, 4, 1), (64, 1024, 3, 2), (48, 200, 5, 1), (33, 77, 4, 3)])
def test_kernel_matches_oracle_on_random_batches(H, W, B, it):
    aug = import_module("htr-vt_b200.augment")
    rs = np.random.RandomState(H + W + B)
    args = _args((8.0 if H >= 48 else 4.0, 3, it, 0.4, 0.4, 0.4, 0.2))
    x = _lines(rs, B, H, W)
    xd = torch.from_numpy(x).cuda()
    seen = set()
    for seed in range(40):
        np.random.seed(1000 + seed); torch.manual_seed(1000 + seed)
        p = aug.draw_collate_params(B, H, W, args)
        combo = tuple(p[k] is not None for k in ("warp", "morph", "jitter"))
        if combo in seen and seed > 12:
            continue
        seen.add(combo)
        got = aug.augment_lines(xd, p).cpu().numpy()
        want = A.apply_params(x, p)
        assert np.array_equal(got, want), (seed, combo, int((got != want).sum()))
    assert len(seen) == 8


def test_anti_aliased_warp_and_strided_input():
    """A quad whose bounding box is taller than 1.25 H: skimage.resize's Gaussian is not the identity any more (three
    taps down the rows).  Also: a [B, 1, H, W] view and a batch slice with a non-trivial image stride."""
    aug = import_module("htr-vt_b200.augment")
    rs = np.random.RandomState(3)
    H, W, B = 64, 320, 3
    x = _lines(rs, B + 1, H, W)
    warp = []
    for b in range(B):
        src = np.array(((-3.0 + b, -8.0), (2.0, H + 8.5 + b), (W + 4.0, H + 7.0), (W - 2.0, -7.5)))
        M, shp = A.projective_from_quad(src, W, H)
        assert shp[0] > 1.25 * H
        warp.append((M, shp))
    p = {"warp": warp, "morph": (2, 3, 1, True), "jitter": [([1, 0, 2, 3], [1.3, 0.7, 1.0, 0.0])] * B}
    rec, _ = aug.pack_params(p, B, H, W)
    assert (rec.view(aug._REC)["gauss"] == 1).all()
    xd = torch.from_numpy(x).cuda()
    got = aug.augment_lines(xd[1:].unsqueeze(1), p)
    assert got.shape == (B, 1, H, W)
    assert np.array_equal(got[:, 0].cpu().numpy(), A.apply_params(x[1:], p))
    # no gate fired: the collate returns the loader's pixels untouched
    none = {"warp": None, "morph": None, "jitter": None}
    assert np.array_equal(aug.augment_lines(xd, none).cpu().numpy(), x)


def test_augmented_uint8_batch_feeds_the_model():
    h = import_module("htrvt_b200")
    Hm = import_module("htr-vt_b200.model.HTR_VT")
    aug = import_module("htr-vt_b200.augment")
    from functools import partial
    rs = np.random.RandomState(4)
    m = Hm.MaskedAutoencoderViT(24, img_size=[64, 128], patch_size=(4, 64), embed_dim=256, depth=1, num_heads=2,
                                mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).cuda().train()
    x = _lines(rs, 4, 64, 128).astype(np.float32) / 255.0
    batch = [(x[i][None], "ab") for i in range(4)]
    np.random.seed(7); torch.manual_seed(7)
    img, labels = aug.SameTrCollate(batch, _args(), as_uint8=True)
    y = m(img, 0.4, 8, use_masking=True)
    assert y.shape == (4, 32, 24) and torch.isfinite(y.float()).all()
    tl = torch.tensor([2, 2, 2, 2], dtype=torch.int32)
    tg = torch.tensor([1, 2] * 4, dtype=torch.int32).cuda()
    h.ctc_loss_from_logits(y, tg, tl).mean().backward()
    assert m.head.weight.grad is not None
