"""CPU oracle of the training-time augmentation `SameTrCollate` (reference model_v1/data/dataset.py:13-45 with
model_v1/data/transform.py:11-33, 164-230 and torchvision's ColorJitter on PIL 'L' images).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg may import this; the product,
htr-vt_b200/, never does).

What the reference runs per batch, on the host, image by image:
  gate 1 (np.random.rand() < 0.5): transform.RandomTransform(args.proj) - a random projective warp
         (skimage.transform.warp, bilinear, cval 255) to the bounding box of the warped corners, then
         skimage.transform.resize back to (h, w), `.astype(np.uint8)` (truncation);
  gate 2: cv2.erode or cv2.dilate with np.ones((kernel_w, kernel_h)) - NB the tuple is a numpy SHAPE, so `kernel_w`
         is the vertical extent - `args.dila_ero_iter` iterations, same kernel for the whole batch;
  gate 3: torchvision ColorJitter(brightness, contrast, saturation, hue) per image; on mode-'L' images saturation and
         hue are identities, brightness / contrast are PIL Image.blend against black / the rounded image mean.
Third-party arithmetic (not under /root/reference), pinned in environment.yaml: opencv-python-headless 4.1.2.30,
pillow 10.3.0, torchvision 0.14.0, scikit-image 0.21.0, numpy 1.24.4.

Pinning.  cv2 4.13, PIL 12.2, torchvision 0.26 and scipy 1.18 are importable in the dev container: `morph`,
`pil_blend` / `jitter_L` and `zoom_linear_mirror` are pinned bit for bit against those libraries
(tests/test_oracle_augment.py) and the committed goldens (tests/golden/augment_cases.npz) are produced by the
reference's OWN SameTrCollate / transform.py source, exec'd unmodified, with the real cv2 / PIL / torchvision.
scikit-image is NOT importable here: **the projective warp + resize is parity unpinned** - `warp_projective`,
`resize_like_skimage`, `homography` restate scikit-image 0.21.0's published algorithm
(skimage/transform/_warps_cy.pyx `_warp_fast`, skimage/_shared/interpolation.pxd `bilinear_interpolation`,
skimage/transform/_warps.py `resize` -> scipy.ndimage.gaussian_filter + scipy.ndimage.zoom(order=1, mode='mirror',
grid_mode=True), skimage/transform/_geometric.py `ProjectiveTransform.estimate`) and stand in for the missing package
(oracle/skimage_stub) when the reference source is exec'd.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------------------------
# gate 2: cv2.erode / cv2.dilate with a rectangular all-ones kernel (transform.py:11-33)
# ------------------------------------------------------------------------------------------------------------------
def morph(x: np.ndarray, k_rows: int, k_cols: int, iterations: int, erode: bool) -> np.ndarray:
    """cv2.erode / cv2.dilate(x, np.ones((k_rows, k_cols)), iterations=it): anchor = kernel centre (size // 2), pixels
    outside the image never win (cv2's default morphology border), `iterations` of a rectangle = one pass with the
    rectangle grown to size + (it - 1)(size - 1) and the anchor scaled by it."""
    H, W = x.shape
    ay, ax = k_rows // 2, k_cols // 2
    lo_y, hi_y = -ay * iterations, -ay * iterations + (k_rows - 1) * iterations
    lo_x, hi_x = -ax * iterations, -ax * iterations + (k_cols - 1) * iterations
    pad = 255 if erode else 0
    m = max(abs(lo_y), abs(hi_y), abs(lo_x), abs(hi_x)) + 1
    xp = np.full((H + 2 * m, W + 2 * m), pad, dtype=np.uint8)
    xp[m:m + H, m:m + W] = x
    acc = None
    for dy in range(lo_y, hi_y + 1):
        for dx in range(lo_x, hi_x + 1):
            s = xp[m + dy:m + dy + H, m + dx:m + dx + W]
            acc = s.copy() if acc is None else (np.minimum(acc, s) if erode else np.maximum(acc, s))
    return acc


# ------------------------------------------------------------------------------------------------------------------
# gate 3: ColorJitter on a mode-'L' PIL image (dataset.py:35-37)
# ------------------------------------------------------------------------------------------------------------------
def pil_blend(degenerate: np.ndarray, px: np.ndarray, f: float) -> np.ndarray:
    """PIL Image.blend(degenerate, image, f) on uint8 (libImaging/Blend.c): float32 arithmetic; inside [0, 1] the result
    is truncated, outside it is clipped to [0, 255] and truncated."""
    f32 = np.float32
    d = degenerate.astype(np.int32)
    t = (d.astype(f32) + f32(f) * (px.astype(np.int32) - d).astype(f32)).astype(f32)
    if f32(0.0) <= f32(f) <= f32(1.0):                          # `float alpha` in C: the test sees the rounded factor
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def jitter_L(x: np.ndarray, order, factors) -> np.ndarray:
    """torchvision ColorJitter.forward on an 'L' image: ops in `order` (a permutation of 0 brightness, 1 contrast,
    2 saturation, 3 hue) with `factors[op]`; 2 and 3 leave 'L' images untouched (ImageEnhance.Color blends the image
    with its own grey version, adjust_hue returns 'L' inputs as they are)."""
    for op in order:
        f = factors[int(op)]
        if f is None:
            continue
        if op == 0:
            x = pil_blend(np.zeros_like(x), x, f)
        elif op == 1:
            mean = int(x.astype(np.float64).sum() / x.size + 0.5)        # int(ImageStat.Stat(img).mean[0] + 0.5)
            x = pil_blend(np.full_like(x, mean), x, f)
    return x


# ------------------------------------------------------------------------------------------------------------------
# gate 1: RandomTransform (transform.py:164-230) - scikit-image 0.21.0 restated, parity unpinned (see header)
# ------------------------------------------------------------------------------------------------------------------
def homography(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """3x3 H (H[2,2] = 1) with dst ~ H src for four point pairs: the unique solution ProjectiveTransform.estimate
    finds by a normalised SVD; here the 8x8 linear system in float64."""
    A, b = [], []
    for (x, y), (u, v) in zip(np.asarray(src, dtype=np.float64), np.asarray(dst, dtype=np.float64)):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y]); b.append(u)
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y]); b.append(v)
    h = np.linalg.solve(np.array(A, dtype=np.float64), np.array(b, dtype=np.float64))
    return np.append(h, 1.0).reshape(3, 3)


def apply_h(H: np.ndarray, pts: np.ndarray) -> np.ndarray:
    p = np.c_[np.asarray(pts, dtype=np.float64), np.ones(len(pts))] @ H.T
    return p[:, :2] / p[:, 2:3]


def warp_projective(img: np.ndarray, M: np.ndarray, out_shape, cval: float) -> np.ndarray:
    """skimage.transform.warp(img, ProjectiveTransform(M), output_shape, order=1, mode='constant', cval,
    preserve_range=True) = `_warp_fast`: output pixel (r, c) samples the input at M (c, r, 1) by bilinear interpolation
    (floor / ceil neighbours, cval outside), float64, in _warp_fast's operation order."""
    img = img.astype(np.float64)
    rows, cols = img.shape
    R, C = int(out_shape[0]), int(out_shape[1])
    y, x = np.meshgrid(np.arange(R, dtype=np.float64), np.arange(C, dtype=np.float64), indexing="ij")
    m = [float(v) for v in np.asarray(M, dtype=np.float64).reshape(-1)]
    xx = m[0] * x + m[1] * y + m[2]
    yy = m[3] * x + m[4] * y + m[5]
    zz = m[6] * x + m[7] * y + m[8]
    c = xx / zz
    r = yy / zz
    minr, minc, maxr, maxc = np.floor(r), np.floor(c), np.ceil(r), np.ceil(c)
    dr, dc = r - minr, c - minc

    def px(rr, cc):
        ok = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
        ri = np.clip(rr, 0, rows - 1).astype(np.int64)
        ci = np.clip(cc, 0, cols - 1).astype(np.int64)
        return np.where(ok, img[ri, ci], float(cval))

    top = (1 - dc) * px(minr, minc) + dc * px(minr, maxc)
    bottom = (1 - dc) * px(maxr, minc) + dc * px(maxr, maxc)
    return (1 - dr) * top + dr * bottom


def _mirror(i: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i < n, i, p - i)


def zoom_linear_mirror(img: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """scipy.ndimage.zoom(img, (oh / ih, ow / iw), order=1, mode='mirror', grid_mode=True) restated (pinned bit for bit
    to scipy in tests/test_oracle_augment.py): input coordinate (o + 0.5) * in / out - 0.5, linear weights (1 - y, y) on
    floor / floor + 1 (of the raw coordinate, also when it is negative), tap indices mirrored about the edge pixel
    centres, sum ((v * wr) * wc) in row-major tap order."""
    img = img.astype(np.float64)
    ih, iw = img.shape
    cr = (np.arange(oh, dtype=np.float64) + 0.5) * (ih / oh) - 0.5
    cc = (np.arange(ow, dtype=np.float64) + 0.5) * (iw / ow) - 0.5
    if ih == 1:                                               # a one-pixel axis: the coordinate is mapped onto pixel 0
        cr = np.zeros(oh)
    if iw == 1:
        cc = np.zeros(ow)
    fr, fc = np.floor(cr), np.floor(cc)
    yr, yc = (cr - fr)[:, None], (cc - fc)[None, :]
    r0, r1 = _mirror(fr.astype(np.int64), ih), _mirror(fr.astype(np.int64) + 1, ih)
    c0, c1 = _mirror(fc.astype(np.int64), iw), _mirror(fc.astype(np.int64) + 1, iw)
    wr0, wr1, wc0, wc1 = 1 - yr, yr, 1 - yc, yc
    # a growing axis starts at a coordinate in (-0.5, 0): floor = -1, and scipy then visits the in-image tap (index
    # 0) BEFORE the mirrored one - the order matters for the last bit of the four-term sum (found by comparing
    # with scipy on fractional data; pinned in tests/test_oracle_augment.py)
    top = (fr < 0)
    r0, r1 = np.where(top, r1, r0), np.where(top, r0, r1)
    wr0, wr1 = np.where(top[:, None], wr1, wr0), np.where(top[:, None], wr0, wr1)
    left = (fc < 0)
    c0, c1 = np.where(left, c1, c0), np.where(left, c0, c1)
    wc0, wc1 = np.where(left[None, :], wc1, wc0), np.where(left[None, :], wc0, wc1)
    t = np.zeros((oh, ow))
    for ri, wr in ((r0, wr0), (r1, wr1)):
        for ci, wc in ((c0, wc0), (c1, wc1)):
            t = t + (img[ri[:, None], ci[None, :]] * wr) * wc
    return t


def antialias_sigma(in_shape, out_shape):
    f = np.divide(np.asarray(in_shape, dtype=np.float64), np.asarray(out_shape, dtype=np.float64))
    return np.maximum(0, (f - 1) / 2)


def resize_like_skimage(img: np.ndarray, out_shape) -> np.ndarray:
    """skimage.transform.resize(img, out_shape, preserve_range=True) of 0.21.0 for a float image: order 1, mode
    'reflect' (= scipy 'mirror'), anti-aliasing Gaussian (sigma = max(0, (in / out - 1) / 2) per axis) whenever an axis
    shrinks, scipy.ndimage.zoom with grid_mode=True, clipped to the input's value range."""
    import scipy.ndimage as ndi
    img = img.astype(np.float64)
    in_shape = img.shape
    oh, ow = int(out_shape[0]), int(out_shape[1])
    filtered = img
    if any(o < i for o, i in zip((oh, ow), in_shape)):
        filtered = ndi.gaussian_filter(img, antialias_sigma(in_shape, (oh, ow)), cval=0, mode="mirror")
    f = np.divide(in_shape, (oh, ow))
    out = ndi.zoom(filtered, [1 / v for v in f], order=1, mode="mirror", cval=0, grid_mode=True)
    return np.clip(out, img.min(), img.max())


def random_transform_params(w: int, h: int, val: float, rng=np.random):
    """The draws of RandomTransform.__call__ (transform.py:176-197) in order -> (M 3x3 inverse map, output_shape)."""
    dw, dh = (val, 0) if rng.randint(0, 2) == 0 else (0, val)

    def rd(d):
        return rng.uniform(-d, d)

    def fd(d):
        return rng.uniform(-dw, d)

    tl_top = rd(dh); tl_left = fd(dw); bl_bottom = rd(dh); bl_left = fd(dw)
    tr_top = rd(dh); tr_right = fd(min(w * 3 / 4 - tl_left, dw))
    br_bottom = rd(dh); br_right = fd(min(w * 3 / 4 - bl_left, dw))
    src = np.array(((tl_left, tl_top), (bl_left, h - bl_bottom), (w - br_right, h - br_bottom), (w - tr_right, tr_top)))
    return projective_from_quad(src, w, h)


def projective_from_quad(src: np.ndarray, w: int, h: int):
    """transform.py:199-226 after the draws: estimate src -> image corners, bounding box of the back-projected corners,
    translation folded in, normalised by the last element.  -> (M, (out_rows, out_cols))."""
    dst = np.array(([0, 0], [0, h - 1], [w - 1, h - 1], [w - 1, 0]), dtype=np.float64)
    H = homography(src, dst)
    corners = apply_h(np.linalg.inv(H), dst)
    minc, minr = corners[:, 0].min(), corners[:, 1].min()
    maxc, maxr = corners[:, 0].max(), corners[:, 1].max()
    out_shape = np.around((maxr - minr + 1, maxc - minc + 1))
    T = np.array([[1, 0, minc], [0, 1, minr], [0, 0, 1]], dtype=np.float64)
    M = H @ T                                                   # `tform4 + tform`: the translation is applied first
    M = M / M[2, 2]
    return M, (int(out_shape[0]), int(out_shape[1]))


def random_transform_apply(x: np.ndarray, M: np.ndarray, out_shape) -> np.ndarray:
    h, w = x.shape
    warped = warp_projective(x, M, out_shape, 255.0)
    return resize_like_skimage(warped, (h, w)).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# the whole collate on a uint8 batch, driven by an explicit parameter record (what the CUDA path receives)
# ------------------------------------------------------------------------------------------------------------------
def apply_params(img_u8: np.ndarray, params: dict) -> np.ndarray:
    """img_u8 [B, H, W]; params as produced by htr-vt_b200/augment.py::draw_collate_params:
    'warp': None | list of (M, out_shape) per image; 'morph': None | (k_rows, k_cols, iterations, erode);
    'jitter': None | list of (order, [b, c, s, h]) per image."""
    out = []
    for i, x in enumerate(img_u8):
        if params.get("warp") is not None:
            M, shp = params["warp"][i]
            x = random_transform_apply(x, np.asarray(M), shp)
        if params.get("morph") is not None:
            kr, kc, it, er = params["morph"]
            x = morph(x, kr, kc, it, er)
        if params.get("jitter") is not None:
            order, fac = params["jitter"][i]
            x = jitter_L(x, order, fac)
        out.append(x)
    return np.stack(out)
