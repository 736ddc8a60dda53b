"""Generate tests/golden/*.npz from the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py
Imports /root/reference/model_v1 and model_window through oracle/refload.py, loads weights made by
htrvt_oracle.init_state_dict (numpy RandomState => reproducible on any box) with
load_state_dict(strict=True), runs the reference's own forward / CTCLoss / decode on seeded
inputs and stores ONLY the outputs (+ the seeds / shapes needed to regenerate the inputs).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import htrvt_oracle as O  # noqa: E402
import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def images(seed, B, W):
    return torch.from_numpy(np.random.RandomState(seed).rand(B, 1, 64, W).astype(np.float32))


def labels(seed, B, C, lo, hi):
    rs = np.random.RandomState(seed)
    lens = rs.randint(lo, hi + 1, size=B).astype(np.int32)
    tg = rs.randint(1, C, size=int(lens.sum())).astype(np.int32)
    return torch.from_numpy(tg), torch.from_numpy(lens)


def grad_digest(grads):
    names = sorted(grads)
    norms = np.array([float(grads[k].double().norm()) for k in names])
    heads = np.stack([np.pad(grads[k].reshape(-1)[:8].numpy(), (0, max(0, 8 - grads[k].numel()))) for k in names])
    return names, norms, heads


def model_case(variant, tag, nb_cls, W, B, seed, cfg, train_seed=7, mask_ratio=0.4, span=8):
    htr, _ = refload.load_variant("model_v1" if variant == "v1" else "model_window")
    kw = dict(nb_cls=nb_cls, img_size=[64, W], patch_size=(4, 64), embed_dim=cfg["embed_dim"],
              depth=cfg["depth"], num_heads=cfg["num_heads"], mlp_ratio=4,
              norm_layer=__import__("functools").partial(torch.nn.LayerNorm, eps=1e-6))
    if variant == "window":
        kw["img_size"] = [W, 64]          # window ctor builds its dummy as [1,1,img_size[1],img_size[0]]
    ref = htr.MaskedAutoencoderViT(**kw)
    sd = O.init_state_dict(nb_cls, [64, W], seed=seed, embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                           num_heads=cfg["num_heads"], variant=variant)
    ref_keys = list(ref.state_dict().keys())
    assert sorted(ref_keys) == sorted(sd.keys()), (set(ref_keys) ^ set(sd.keys()))
    for k in ref_keys:
        assert tuple(ref.state_dict()[k].shape) == tuple(sd[k].shape), k
    ref.load_state_dict(sd, strict=True)
    x = images(seed + 1, B, W)
    out = {"keys": np.array(ref_keys)}
    ref.eval()
    with torch.no_grad():
        out["logits_eval"] = ref(x).numpy()
    if variant == "v1":
        # train mode: span mask drawn from the CPU default generator inside the reference
        ref.train()
        tg, tl = labels(seed + 2, B, nb_cls, 4, 12)
        torch.manual_seed(train_seed)
        preds = ref(x, mask_ratio, span, use_masking=True)
        lp = preds.float().permute(1, 0, 2).log_softmax(2)
        crit = torch.nn.CTCLoss(reduction="none", zero_infinity=True)
        in_len = torch.IntTensor([preds.size(1)] * B)
        loss = crit(lp, tg, in_len, tl).mean()
        loss.backward()
        grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        names, norms, heads = grad_digest(grads)
        out.update(logits_train=preds.detach().numpy(), loss=np.float64(loss.item()),
                   grad_names=np.array(names), grad_norms=norms, grad_heads=heads,
                   bn1_running_mean=ref.state_dict()["patch_embed.bn1.running_mean"].numpy(),
                   l3_running_var=ref.state_dict()["patch_embed.layer3.1.bn2.running_var"].numpy(),
                   nbt=np.int64(ref.state_dict()["patch_embed.bn1.num_batches_tracked"].item()))
        torch.manual_seed(train_seed)
        out["mask"] = O.draw_span_mask(W // 4, mask_ratio, span).numpy()
    out["meta"] = np.array([nb_cls, W, B, seed, cfg["embed_dim"], cfg["depth"], cfg["num_heads"], train_seed])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag, {k: getattr(v, "shape", None) for k, v in out.items()})


grad_sample_index = O.grad_sample_index


def train_batch_case(tag="v1_train_b32", nb_cls=80, W=512, B=32, seed=777, train_seed=9):
    """model_v1, FULL architecture, a batch where BatchNorm batch statistics are well conditioned (B = 32):
    train-mode logits / loss / per-sample nll / gradient samples of all 101 trainable tensors through the reference's
    own compute_loss sequence (model_v1/train.py:21-30), plus eval-mode logits and the strings the reference's
    valid.py:40-42 + CTCLabelConverter.decode produce from them."""
    htr, utl = refload.load_variant("model_v1")
    ref = htr.create_model(nb_cls=nb_cls, img_size=[64, W])
    sd = O.init_state_dict(nb_cls, [64, W], seed=seed)
    ref.load_state_dict(sd, strict=True)
    x = images(seed + 1, B, W)
    tg, tl = labels(seed + 2, B, nb_cls, 16, 64)
    out = {"meta": np.array([nb_cls, W, B, seed, train_seed])}
    ref.eval()
    alphabet = "".join(chr(33 + i) for i in range(nb_cls - 1))
    conv = utl.CTCLabelConverter(alphabet)
    with torch.no_grad():
        pe = ref(x).float()
        out["logits_eval"] = pe.numpy()
        preds = pe.permute(1, 0, 2).log_softmax(2)                      # valid.py:32,35
        _, idx = preds.max(2)                                           # valid.py:40
        idx = idx.transpose(1, 0).contiguous().view(-1)                 # valid.py:41
        out["strings_eval"] = np.array(conv.decode(idx.data, torch.IntTensor([pe.size(1)] * B)))   # valid.py:42
        out["index_eval"] = idx.numpy().astype(np.int16)
    ref.train()
    torch.manual_seed(train_seed)
    preds = ref(x, 0.4, 8, use_masking=True)
    lp = preds.float().permute(1, 0, 2).log_softmax(2)
    nll = torch.nn.CTCLoss(reduction="none", zero_infinity=True)(lp, tg, torch.IntTensor([preds.size(1)] * B), tl)
    loss = nll.mean()
    loss.backward()
    torch.manual_seed(train_seed)
    out["mask"] = O.draw_span_mask(W // 4, 0.4, 8).numpy()
    out.update(logits_train=preds.detach().numpy(), loss=np.float64(loss.item()), nll=nll.detach().numpy())
    names = [k for k, p in ref.named_parameters() if p.grad is not None]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([float(dict(ref.named_parameters())[k].grad.double().norm()) for k in names])
    params = dict(ref.named_parameters())
    out["grad_samples"] = np.concatenate(
        [params[k].grad.reshape(-1)[torch.from_numpy(grad_sample_index(k, params[k].numel()))].numpy() for k in names])
    out["grad_sample_counts"] = np.array([len(grad_sample_index(k, params[k].numel())) for k in names])
    msd = ref.state_dict()
    out["bn_running"] = np.concatenate([msd[k].numpy().reshape(-1) for k in msd if "running_" in k])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print("wrote", tag, float(loss), out["strings_eval"][:2])


def window_ragged_case(tag, W, B, seed, nb_cls=90):
    """model_window at a width whose token count is NOT a multiple of the 16-token window: exercises the zero-pad +
    rolled key-padding-mask branch (model_window/model/HTR_VT.py:121-131,49-56).  Eval-mode logits of the reference."""
    htr, _ = refload.load_variant("model_window")
    ref = htr.create_model(nb_cls=nb_cls, img_size=[W, 64])
    sd = O.init_state_dict(nb_cls, [64, W], seed=seed, variant="window")
    assert sorted(ref.state_dict().keys()) == sorted(sd.keys())
    ref.load_state_dict(sd, strict=True)
    x = images(seed + 1, B, W)
    ref.eval()
    with torch.no_grad():
        lg = ref(x).numpy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), logits_eval=lg, meta=np.array([nb_cls, W, B, seed]),
                        keys=np.array(list(ref.state_dict().keys())))
    print("wrote", tag, lg.shape)


def ctc_cases():
    """torch.nn.CTCLoss (the reference's criterion, model_v1/train.py:95) on CPU."""
    specs = [  # name, B, T, C, (Lmin, Lmax), special
        ("basic", 4, 32, 12, (3, 10), None),
        ("iam_shape", 6, 128, 80, (16, 64), None),
        ("repeats", 4, 24, 5, (6, 11), "repeats"),
        ("infeasible", 4, 10, 6, (2, 9), "infeasible"),
        ("empty_target", 3, 12, 7, (0, 3), "empty"),
        ("single_state", 2, 1, 4, (1, 1), None),
        ("long_labels", 3, 256, 90, (150, 200), None),
        ("ragged_T", 4, 40, 10, (3, 9), "ragged"),
    ]
    out = {}
    for name, B, T, C, (lo, hi), special in specs:
        rs = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        logits = (rs.randn(B, T, C) * 2.0).astype(np.float32)
        lens = rs.randint(lo, hi + 1, size=B).astype(np.int32)
        if special == "empty":
            lens[0] = 0
        if special == "infeasible":
            lens[1] = 9          # needs >= 9 frames (+repeats) out of 10: likely infeasible with repeats
        tg = rs.randint(1, C, size=int(lens.sum())).astype(np.int32)
        if special in ("repeats", "infeasible"):
            tg[1::2] = tg[0::2][: len(tg[1::2])]       # force adjacent repeated labels
        in_len = np.full(B, T, dtype=np.int32)
        if special == "ragged":
            in_len = rs.randint(20, T + 1, size=B).astype(np.int32)
        lg = torch.from_numpy(logits).requires_grad_(True)
        lp = lg.permute(1, 0, 2).log_softmax(2)
        nll = torch.nn.CTCLoss(reduction="none", zero_infinity=True)(
            lp, torch.from_numpy(tg), torch.from_numpy(in_len), torch.from_numpy(lens))
        nll.sum().backward()
        out[name + ".logits"] = logits
        out[name + ".targets"] = tg
        out[name + ".in_len"] = in_len
        out[name + ".tgt_len"] = lens
        out[name + ".nll"] = nll.detach().numpy()
        out[name + ".grad"] = lg.grad.numpy()
        print("ctc", name, nll.detach().numpy()[:4])
    out["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(os.path.join(OUT, "ctc_cases.npz"), **out)


def decode_cases():
    """CTCLabelConverter.decode (model_v1/utils/utils.py:72-86) on adversarial index streams."""
    _, utl = refload.load_variant("model_v1")
    alphabet = "abcdefghijklmnopqrstuvwxyz0123456789 .,'-"
    conv = utl.CTCLabelConverter(alphabet)
    rs = np.random.RandomState(5)
    T = 24
    rows = [
        np.zeros(T, dtype=np.int64),                                   # all blank
        np.full(T, 3, dtype=np.int64),                                 # one long repeat
        np.tile([1, 0, 1, 1, 0, 0, 2, 2], 3),                          # repeats across blanks
        np.tile([5, 60, 5, 5, 60, 60, 7, 0], 3),                       # ids >= len(character) dropped, still "previous"
        rs.randint(0, len(alphabet) + 1, size=T),
        rs.randint(0, 6, size=T),
        np.arange(T) % (len(alphabet) + 6),
        np.array([41] * 3 + [42] * 3 + [43] * 3 + [0] * 3 + [41, 42] * 6),
    ]
    idx = np.concatenate(rows).astype(np.int64)
    lens = np.full(len(rows), T, dtype=np.int32)
    strs = conv.decode(torch.from_numpy(idx), torch.from_numpy(lens))
    # ragged lengths too
    lens2 = np.array([5, 24, 1, 30, 24, 12, 48, 24, 24], dtype=np.int32)
    assert lens2.sum() == idx.size
    strs2 = conv.decode(torch.from_numpy(idx), torch.from_numpy(lens2))
    # READ2016 special case: 87-char alphabet gets '[' -> 88, ']' -> 89 (utils.py:61-62)
    conv87 = utl.CTCLabelConverter("".join(chr(48 + i) for i in range(87)))
    enc87 = conv87.encode(["0a", "[x]"])
    np.savez_compressed(os.path.join(OUT, "decode_cases.npz"), alphabet=np.array(alphabet), index=idx,
                        lens=lens, strings=np.array(strs), lens2=lens2, strings2=np.array(strs2),
                        enc87_text=enc87[0].cpu().numpy(), enc87_len=enc87[1].cpu().numpy(),
                        n_character87=np.int64(len(conv87.character)))
    print("decode", strs, strs2)


def argmax_cases():
    """torch.max(2) index semantics (model_v1/valid.py:40): ties and NaNs."""
    a = np.random.RandomState(9).randn(3, 6, 7).astype(np.float32)
    a[0, 0, 2] = a[0, 0, 5] = 9.0
    a[0, 1, :] = 1.0
    a[1, 2, 4] = np.nan
    a[1, 3, 1] = np.nan
    a[1, 3, 6] = np.nan
    a[2, 0, 3] = np.inf
    a[2, 1, :] = -np.inf
    _, idx = torch.from_numpy(a).max(2)
    np.savez_compressed(os.path.join(OUT, "argmax_cases.npz"), logits=a, index=idx.numpy())
    print("argmax", idx.numpy().tolist())


def metrics_cases():
    """format_string_for_wer comes from the reference itself (model_v1/utils/utils.py:176-179, imported unmodified);
    `editdistance` (environment.yaml:35) is not installed here, so the distances are the oracle's restatement of the
    recurrence the reference spells out at model_v1/test.py:114-133 plus classic known answers."""
    import json
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import htrvt_oracle as O
    _, utl = refload.load_variant("model_v1")
    texts = ["Hello, world!  it's\u2014a [test]\n ok", "a-b_c \u20ac5 #1 100% 3\u00b0 x\\y \"q\" {z}/(w)&+*=<>?;:", "", "   ",
             "no punctuation here", "tab\there", "He said: (well...) nothing-much_at all!", "MOVE to stop Mr. Gaitskell from"]
    fmt = [utl.format_string_for_wer(t) for t in texts]
    kat = [["kitten", "sitting", 3], ["flaw", "lawn", 2], ["", "abc", 3], ["abc", "", 3], ["same", "same", 0],
           ["intention", "execution", 5], ["sunday", "saturday", 3], ["a", "b", 1]]
    preds = ["MOVE to stop Mr. Gaitskell from", "nominating any more Labour life Peers", "", "x", "is to be made at a meeting"]
    labels = ["A MOVE to stop Mr. Gaitskell from", "nominating any more Labour life Peers", "abc", "", "is to be made at a meeting of Labour"]
    rates = O.error_rates(preds, labels)
    with open(os.path.join(OUT, "metrics_cases.json"), "w") as fh:
        json.dump(dict(texts=texts, formatted=fmt, kat=kat, preds=preds, labels=labels, rates=rates), fh, indent=1)
    print("metrics", rates)


def line_prep_cases():
    """u8 line batch -> what the reference's loader + first LayerNorm produce (dataset.py:44,129-130; HTR_VT.py:224),
    computed with torch exactly as the reference does (float() / 255., pad 1.0, F.layer_norm)."""
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, size=(3, 64, 128)).astype(np.uint8)
    widths = np.array([128, 77, 4], dtype=np.int32)
    x = torch.from_numpy(img).float() / 255.
    for b, w in enumerate(widths):
        x[b, :, int(w):] = 1.0
    y = torch.nn.functional.layer_norm(x, x.shape[1:], eps=1e-5)
    np.savez_compressed(os.path.join(OUT, "line_prep_cases.npz"), img=img, widths=widths, y=y.numpy())
    print("line_prep", float(y.abs().max()))


class StubLM(object):
    """Deterministic stand-in for KenLMTextScorer (kenlm is not installed here): any object with .score(text)."""

    def score(self, text):
        return sum(((ord(ch) * 31 + i * 17) % 97) / 97.0 for i, ch in enumerate(text)) - 0.6 * len(text)


def reference_beam_function():
    """The reference's `simple_ctc_beam_search_with_lm`, taken from its source file UNMODIFIED (the module itself
    cannot be imported: it needs kenlm at import time) and exec'd with numpy in scope."""
    import ast
    path = os.path.join(refload.REF_ROOT, "model_window", "test_with_kenlm.py")
    src = open(path).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "simple_ctc_beam_search_with_lm"][0]
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["simple_ctc_beam_search_with_lm"]


def beam_cases():
    """Best strings of the reference's LM-rescored per-frame beam (model_window/test_with_kenlm.py:25-59) on random
    and peaked log-prob lines, with the reference's own CTCLabelConverter.  The reference accumulates path scores as
    `0.0 + np.float32` - float64 under its pinned numpy 1.24, float32 under the numpy 2 of this container; the cases
    are kept only if the oracle's float64 restatement and a float32 variant agree on every candidate list, so the
    fixtures do not depend on the promotion rule."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import htrvt_oracle as O
    _, utl = refload.load_variant("model_window")
    beam_fn = reference_beam_function()
    alphabet = "abcdefghijklmnopqrstuvwxyz .,'"
    conv = utl.CTCLabelConverter(alphabet)
    lm = StubLM()
    rs = np.random.RandomState(17)
    T, C = 40, len(alphabet) + 1
    lines, best, ks = [], [], []
    tries = 0
    while len(lines) < 12 and tries < 200:
        tries += 1
        kind = len(lines) % 3
        x = rs.randn(T, C).astype(np.float32) * (1.0, 3.0, 6.0)[kind]
        if kind == 2:
            x[:, 0] += 4.0                                     # blank-heavy, as a trained network emits
        lp = torch.from_numpy(x).log_softmax(1)
        K = (5, 3, 8)[kind]
        c64 = O.kbest_paths(lp.numpy(), K, np.float64)
        c32 = O.kbest_paths(lp.numpy(), K, np.float32)
        if [c[0] for c in c64] != [c[0] for c in c32]:
            continue
        s = beam_fn(lp, conv, lm, beam_size=K)
        lines.append(lp.numpy()); best.append(s); ks.append(K)
    np.savez_compressed(os.path.join(OUT, "beam_cases.npz"), alphabet=np.array(alphabet), log_probs=np.stack(lines),
                        beam=np.array(ks, dtype=np.int32), best=np.array(best))
    print("beam", len(lines), "cases;", tries, "draws;", best[:4])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "r2":           # the round-2 additions only
        window_ragged_case("win_w1000", 1000, 2, 411)
        window_ragged_case("win_w600", 600, 2, 412)
        train_batch_case()
        sys.exit(0)
    small = dict(embed_dim=64, depth=2, num_heads=2)
    full = dict(embed_dim=768, depth=4, num_heads=6)
    model_case("v1", "v1_small", 20, 128, 3, 11, small)
    model_case("v1", "v1_full", 80, 512, 2, 123, full)
    model_case("window", "win_small", 20, 128, 3, 21, dict(embed_dim=64, depth=4, num_heads=2))
    model_case("window", "win_full", 90, 1024, 1, 321, full)
    ctc_cases()
    decode_cases()
    argmax_cases()
    metrics_cases()
    line_prep_cases()
    beam_cases()
    window_ragged_case("win_w1000", 1000, 2, 411)
    window_ragged_case("win_w600", 600, 2, 412)
    train_batch_case()
