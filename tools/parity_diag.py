"""Print (not assert) the parity numbers of the drop-in path on the GPU box: train-mode logits error vs the B = 32
reference golden, per-family gradient cosines, eval-mode error and decode agreement.  Developer aid for tightening
the tolerances in tests/test_gpu_model.py."""
import os
import sys
from importlib import import_module

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import htrvt_oracle as O  # noqa: E402
import test_gpu_model as T  # noqa: E402


def main():
    import htrvt_b200 as h
    g = np.load(os.path.join(ROOT, "tests", "golden", "v1_train_b32.npz"))
    nb_cls, W, B, seed, train_seed = [int(v) for v in g["meta"]]
    m, sd = T._build(nb_cls, W, 768, 4, 6, seed)
    x = T._images(seed + 1, B, W)
    tg, tl = T._labels(seed + 2, B, nb_cls, 16, 64)
    prec = os.environ.get("HTRVT_DIAG_PRECISION")
    if prec and hasattr(m, "set_precision"):
        m.set_precision(prec)
    m.eval()
    with torch.no_grad():
        pe = m(x.cuda()).float().cpu().numpy()
    print("eval  relerr vs reference golden: %.3e" % T._relerr(pe, g["logits_eval"]))
    idx = pe.argmax(-1).reshape(-1)
    print("eval  argmax mismatches: %d of %d" % (int((idx != g["index_eval"].astype(np.int64)).sum()), idx.size))
    if prec == "fp32":
        return
    m.train()
    torch.manual_seed(train_seed)
    preds = m(x.cuda(), 0.4, 8, use_masking=True)
    loss = h.ctc_loss_from_logits(preds.float(), tg.cuda(), tl).mean()
    loss.backward()
    print("train relerr vs reference golden: %.3e  loss %.4f vs %.4f" %
          (T._relerr(preds.detach().float().cpu().numpy(), g["logits_train"]), loss.item(), float(g["loss"])))
    offs = np.concatenate([[0], np.cumsum(g["grad_sample_counts"])])
    worst = {"stem": (2.0, ""), "transformer": (2.0, "")}
    rows = []
    for i, name in enumerate(g["grad_names"].tolist()):
        a = dict(m.named_parameters())[name].grad.detach().float().cpu()
        idx = torch.from_numpy(O.grad_sample_index(name, a.numel()))
        a_s = a.reshape(-1)[idx].double()
        b_s = torch.from_numpy(g["grad_samples"][offs[i]:offs[i + 1]]).double()
        cos = float((a_s @ b_s) / (a_s.norm() * b_s.norm() + 1e-30))
        ratio = float(a.double().norm() / (float(g["grad_norms"][i]) + 1e-30))
        rows.append((cos, ratio, name))
        fam = T._family(name)
        if cos < worst[fam][0]:
            worst[fam] = (cos, name)
    rows.sort()
    for cos, ratio, name in rows[:12]:
        print("  %-50s cos %.5f ratio %.4f" % (name, cos, ratio))
    print("worst per family:", worst)
    print("ratio range: %.4f .. %.4f" % (min(r[1] for r in rows), max(r[1] for r in rows)))


if __name__ == "__main__":
    main()
