"""Development diagnostic (GPU box): per-tensor parity of one train step against the CPU oracle."""
import os
import sys
from functools import partial
from importlib import import_module

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import htrvt_b200 as h  # noqa: E402
import htrvt_oracle as O  # noqa: E402

H = import_module("htr-vt_b200.model.HTR_VT")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30), np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)


def run(nb_cls, W, D, depth, heads, B, seed):
    m = H.MaskedAutoencoderViT(nb_cls, img_size=[64, W], patch_size=(4, 64), embed_dim=D, depth=depth, num_heads=heads,
                               mlp_ratio=4, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    sd = O.init_state_dict(nb_cls, [64, W], seed=seed, embed_dim=D, depth=depth, num_heads=heads)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    rs = np.random.RandomState(seed + 1)
    x = torch.from_numpy(rs.rand(B, 1, 64, W).astype(np.float32))
    tl = torch.from_numpy(rs.randint(4, 13, size=B).astype(np.int32))
    tg = torch.from_numpy(rs.randint(1, nb_cls, size=int(tl.sum())).astype(np.int32))
    torch.manual_seed(7)
    preds = m(x.cuda(), 0.4, 8, use_masking=True)
    loss = h.ctc_loss_from_logits(preds, tg.cuda(), tl).mean()
    loss.backward()
    torch.manual_seed(7)
    mask = O.draw_span_mask(W // 4, 0.4, 8)
    sd_ref = {k: v.clone() for k, v in sd.items()}
    ref_loss, ref_grads, ref_logits = O.train_step(sd_ref, x, tg, tl, mask, num_heads=heads)
    print("config", nb_cls, W, D, depth, heads, B)
    print("logits relmax %.4g relL2 %.4g" % rel(preds.detach().cpu().numpy(), ref_logits.numpy()))
    print("loss %.6f ref %.6f" % (loss.item(), ref_loss))
    # library bf16 path for calibration: the oracle under autocast on the GPU
    sd_gpu = {k: v.clone().cuda() for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ac = O.forward(sd_gpu, x.cuda(), mask=mask.cuda(), training=True, num_heads=heads).float().cpu()
    print("torch-autocast-bf16 logits relmax %.4g relL2 %.4g" % rel(ac.numpy(), ref_logits.numpy()))
    for name, p in m.named_parameters():
        if name == "pos_embed":
            continue
        a = p.grad.detach().float().cpu().double().reshape(-1)
        b = ref_grads[name].double().reshape(-1)
        cos = float((a @ b) / (a.norm() * b.norm() + 1e-30))
        print("%-50s cos %.5f ratio %.4f |ref| %.3e" % (name, cos, float(a.norm() / (b.norm() + 1e-30)), float(b.norm())))


if __name__ == "__main__":
    run(24, 128, 256, 2, 2, 3, 5)
    run(80, 512, 768, 4, 6, 2, 123)
    run(80, 512, 768, 4, 6, 16, 321)
