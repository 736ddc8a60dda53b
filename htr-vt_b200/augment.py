"""Training-time augmentation of a batch of line images on the GPU: the reference's
`SameTrCollate` (model_v1/data/dataset.py:13-45) - random projective warp (transform.RandomTransform,
model_v1/data/transform.py:164-230), erosion / dilation (transform.py:11-33, cv2) and torchvision ColorJitter on
mode-'L' images - as ONE kernel over the uint8 batch (csrc/augment.cu, one CTA per line image, everything in shared
memory), feeding the uint8 input kernel of `model.forward` (csrc/pipeline.cu) without a float round trip.

The random decisions stay on the host and consume numpy's and torch's GLOBAL generators in exactly the reference's
order (`draw_collate_params`), so a seeded run takes the same augmentation decisions as the reference collate; the
device receives one small parameter record per image.  The reference does the pixel work per image in PIL / OpenCV /
scikit-image on the host (~10 ms per image).

    images, labels = SameTrCollate(batch, args)               # reference signature, float [B,1,H,W] in [0,1] (CUDA)
    images_u8, labels = SameTrCollate(batch, args, as_uint8=True)    # uint8 [B,1,H,W] for model.forward(uint8)
"""
import numpy as np
import torch

from . import ops

_GAUSS_TRUNCATE = 4.0            # scipy.ndimage.gaussian_filter default, what skimage.transform.resize uses


def _jitter_range(value, center=1.0, bound=(0.0, float("inf")), clip_first_on_zero=True):
    """torchvision ColorJitter._check_input: the sampling interval of one factor, None = the op is off."""
    if value is None:
        return None
    if isinstance(value, (int, float)):
        if value < 0:
            raise ValueError("jitter strength must be non negative")
        lo, hi = center - float(value), center + float(value)
        if clip_first_on_zero:
            lo = max(lo, 0.0)
    else:
        lo, hi = float(value[0]), float(value[1])
    if not bound[0] <= lo <= hi <= bound[1]:
        raise ValueError("jitter values should be between %s" % (bound,))
    if lo == hi == center:
        return None
    return lo, hi


def _random_quad_draw(w, h, val):
    """transform.RandomTransform.__call__ (transform.py:176-197): the nine draws from numpy's global generator ->
    the four source points (tl, bl, br, tr) that are mapped onto the image corners."""
    rng = np.random
    dw, dh = (val, 0) if rng.randint(0, 2) == 0 else (0, val)
    tl_top = rng.uniform(-dh, dh); tl_left = rng.uniform(-dw, dw)
    bl_bottom = rng.uniform(-dh, dh); bl_left = rng.uniform(-dw, dw)
    tr_top = rng.uniform(-dh, dh); tr_right = rng.uniform(-dw, min(w * 3 / 4 - tl_left, dw))
    br_bottom = rng.uniform(-dh, dh); br_right = rng.uniform(-dw, min(w * 3 / 4 - bl_left, dw))
    return ((tl_left, tl_top), (bl_left, h - bl_bottom), (w - br_right, h - br_bottom), (w - tr_right, tr_top))


def _maps_from_quads(quads, w, h):
    """transform.py:199-226 for a batch of quads [B, 4, 2]: homography quad -> image corners (the 8 x 8 system of
    ProjectiveTransform.estimate, all B systems in one LAPACK call), bounding box of the back-projected corners,
    translation folded in, normalised.  -> [(M 3x3 inverse map, (rows, cols) of the intermediate warped image)] * B."""
    src = np.asarray(quads, dtype=np.float64)
    B = src.shape[0]
    dst = np.array(([0, 0], [0, h - 1], [w - 1, h - 1], [w - 1, 0]), dtype=np.float64)
    x, y = src[:, :, 0], src[:, :, 1]
    u, v = dst[None, :, 0], dst[None, :, 1]
    one, zero = np.ones_like(x), np.zeros_like(x)
    A = np.empty((B, 8, 8))
    A[:, 0::2] = np.stack([x, y, one, zero, zero, zero, -u * x, -u * y], axis=-1)
    A[:, 1::2] = np.stack([zero, zero, zero, x, y, one, -v * x, -v * y], axis=-1)
    rhs = np.empty((B, 8))
    rhs[:, 0::2], rhs[:, 1::2] = u, v
    Hm = np.concatenate([np.linalg.solve(A, rhs[:, :, None])[:, :, 0], np.ones((B, 1))], axis=1).reshape(B, 3, 3)
    p = np.c_[dst, np.ones(4)][None] @ np.transpose(np.linalg.inv(Hm), (0, 2, 1))
    corners = p[:, :, :2] / p[:, :, 2:3]
    minc, minr = corners[:, :, 0].min(1), corners[:, :, 1].min(1)
    maxc, maxr = corners[:, :, 0].max(1), corners[:, :, 1].max(1)
    shapes = np.around(np.stack([maxr - minr + 1, maxc - minc + 1], axis=1))
    T = np.tile(np.eye(3), (B, 1, 1))
    T[:, 0, 2], T[:, 1, 2] = minc, minr
    M = Hm @ T
    M = M / M[:, 2:3, 2:3]
    return [(M[i], (int(shapes[i, 0]), int(shapes[i, 1]))) for i in range(B)]


def draw_collate_params(B, H, W, args):
    """The random decisions of SameTrCollate (dataset.py:21-37) for a batch of B images of H x W, drawn from the global
    numpy / torch generators in the reference's order.  -> {'warp': None | [(M, (rows, cols))] * B,
    'morph': None | (k_rows, k_cols, iterations, erode), 'jitter': None | [(order, [b, c, s, h])] * B}."""
    from torchvision.transforms import ColorJitter
    p = {"warp": None, "morph": None, "jitter": None}
    if np.random.rand() < 0.5:
        p["warp"] = _maps_from_quads([_random_quad_draw(W, H, args.proj) for _ in range(B)], W, H)
    if np.random.rand() < 0.5:
        kernel_h = int(torch.randint(1, args.dila_ero_max_kernel + 1, (1,)))
        kernel_w = int(torch.randint(1, args.dila_ero_max_kernel + 1, (1,)))
        erode = int(torch.randint(0, 2, (1,))) == 0
        # np.ones((kernel_w, kernel_h)): numpy SHAPE order, `kernel_w` is the number of kernel ROWS (transform.py:16,30)
        p["morph"] = (kernel_w, kernel_h, int(args.dila_ero_iter), erode)
    if np.random.rand() < 0.5:
        rb = _jitter_range(args.jitter_brightness)
        rc = _jitter_range(args.jitter_contrast)
        rs = _jitter_range(args.jitter_saturation)
        rh = _jitter_range(args.jitter_hue, center=0.0, bound=(-0.5, 0.5), clip_first_on_zero=False)
        jit = []
        for _ in range(B):
            order, b, c, s, hh = ColorJitter.get_params(rb, rc, rs, rh)
            jit.append(([int(v) for v in order], [b, c, s, hh]))
        p["jitter"] = jit
    return p


_REC = np.dtype([("m", "<f8", (9,)), ("w0", "<f8"), ("w1", "<f8"), ("warp", "<i4"), ("rows", "<i4"), ("cols", "<i4"),
                 ("gauss", "<i4"), ("jit_n", "<i4"), ("jit_op", "<i4", (2,)), ("jit_f", "<f4", (2,)), ("pad", "<i4")])
assert _REC.itemsize == 128


def pack_params(params, B, H, W):
    """-> (uint8 record array [B, 128] for htrvt_augment_lines, morph tuple of 4 ints).  Raises if the warp's
    anti-aliasing filter (skimage.transform.resize, sigma = (in / out - 1) / 2 per shrinking axis) needs more than the
    3 vertical taps the kernel has (never with the reference's proj = 8 on 64-pixel lines)."""
    rec = np.zeros(B, dtype=_REC)
    if params.get("warp") is not None:
        for i, (M, (rows, cols)) in enumerate(params["warp"]):
            rec["m"][i] = np.asarray(M, dtype=np.float64).reshape(-1)
            rec["warp"][i], rec["rows"][i], rec["cols"][i] = 1, rows, cols
            if rows < 1 or cols < 1 or rows > 4 * H or cols > 4 * W:
                raise ops.HtrvtError("augment: degenerate warp output shape %s" % ((rows, cols),))
            sig = np.maximum(0, (np.divide((rows, cols), (H, W)) - 1) / 2) if (H < rows or W < cols) else np.zeros(2)
            rad = [int(_GAUSS_TRUNCATE * float(s) + 0.5) if s > 1e-15 else 0 for s in sig]
            if rad[0] > 1 or rad[1] > 0:
                raise ops.HtrvtError("augment: projection value too large for the fused warp (anti-aliasing radius %s)"
                                     % (rad,))
            if rad[0] == 1:
                x = np.arange(-1, 2)
                phi = np.exp(-0.5 / (float(sig[0]) * float(sig[0])) * x ** 2)
                phi = phi / phi.sum()
                rec["gauss"][i], rec["w0"][i], rec["w1"][i] = 1, phi[1], phi[0]
    if params.get("jitter") is not None:
        for i, (order, fac) in enumerate(params["jitter"]):
            n = 0
            for op in order:
                if op in (0, 1) and fac[op] is not None:       # saturation / hue: identities on grey images
                    rec["jit_op"][i, n], rec["jit_f"][i, n] = op, fac[op]
                    n += 1
            rec["jit_n"][i] = n
    morph = (0, 1, 1, 1)
    if params.get("morph") is not None:
        kr, kc, it, erode = params["morph"]
        morph = (1 if erode else 2, int(kr), int(kc), int(it))
    return rec.view(np.uint8).reshape(B, 128), morph


def augment_lines(img_u8, params):
    """img_u8: CUDA uint8 [B, H, W] (or [B, 1, H, W]); params from draw_collate_params.  -> CUDA uint8, same shape."""
    shape = img_u8.shape
    x = img_u8.reshape(shape[0], shape[-2], shape[-1])
    B, H, W = x.shape
    rec, morph = pack_params(params, B, H, W)
    return ops.augment_lines(x, torch.from_numpy(rec), morph).reshape(shape)


def SameTrCollate(batch, args, device="cuda", as_uint8=False):
    """Reference signature (dataset.py:13): batch = [(image float [1, H, W] in [0, 1], label)], args with proj,
    dila_ero_max_kernel, dila_ero_iter, jitter_{brightness,contrast,saturation,hue}.
    -> (image tensor [B, 1, H, W] on `device`: float32 in [0, 1] like the reference, or uint8 with as_uint8=True -
    what model.forward takes directly -, labels)."""
    images, labels = zip(*batch)
    u8 = np.stack([np.uint8(np.asarray(im)[0] * 255) for im in images])             # dataset.py:16-17
    B, H, W = u8.shape
    params = draw_collate_params(B, H, W, args)
    x = torch.from_numpy(u8).to(device, non_blocking=True)
    if any(params[k] is not None for k in ("warp", "morph", "jitter")):
        x = augment_lines(x, params)
    x = x.unsqueeze(1)
    if not as_uint8:
        # tensor / tensor: an IEEE division like the reference's CPU `image_tensors / 255.` (tensor / python scalar
        # multiplies by a rounded reciprocal on CUDA and is off by an ulp on some grey levels)
        x = x.float() / torch.full((), 255., device=x.device)
    return x, labels
