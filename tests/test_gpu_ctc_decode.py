"""GPU parity: CTC loss/grad and greedy decode kernels (through the C ABI) vs the oracle / goldens."""
import os

import numpy as np
import pytest
import torch

import htrvt_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _pkg():
    import htrvt_b200
    return htrvt_b200


@pytest.fixture(params=["cta_per_sequence", "lane_group"], autouse=True)
def _ctc_kernel(request):
    """Every CTC test runs against BOTH kernels: the CTA-per-sequence one (mode 0) and the lane-group throughput kernel
    forced at any batch size (mode 1; sequences it flags are redone by the first kernel's fix-up launch)."""
    if "ctc" not in request.node.name:
        yield
        return
    from importlib import import_module
    import htrvt_b200  # noqa: F401
    lib = import_module("htr-vt_b200._lib").lib()
    prev = lib.htrvt_ctc_set_mode({"cta_per_sequence": 0, "lane_group": 1}[request.param])
    yield
    lib.htrvt_ctc_set_mode(prev)


def _ctc_ref64(logits, tg, il, tl):
    """float64 torch CPU CTC (same recursion as the float64 numpy oracle it is pinned to in test_oracle)."""
    lg = torch.from_numpy(logits).double().requires_grad_(True)
    lp = lg.permute(1, 0, 2).log_softmax(2)
    nll = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg).long(), torch.from_numpy(il).long(),
                                       torch.from_numpy(tl).long(), blank=0, reduction="none", zero_infinity=True)
    nll.sum().backward()
    return nll.detach().numpy(), lg.grad.numpy()


def test_ctc_golden_cases_fused_logits():
    h = _pkg()
    g = np.load(os.path.join(G, "ctc_cases.npz"))
    for name in g["names"].tolist():
        logits, tg, il, tl = g[name + ".logits"], g[name + ".targets"], g[name + ".in_len"], g[name + ".tgt_len"]
        x = torch.from_numpy(logits).cuda().requires_grad_(True)
        nll = h.ctc_loss_from_logits(x, torch.from_numpy(tg).cuda(), torch.from_numpy(tl).cuda(),
                                     input_lengths=torch.from_numpy(il).cuda())
        nll.sum().backward()
        ref_nll, ref_grad = _ctc_ref64(logits, tg, il, tl)
        # tolerance stated by north_star: 1e-4 relative (fp32)
        np.testing.assert_allclose(nll.detach().cpu().numpy(), ref_nll, rtol=1e-4, atol=1e-4, err_msg=name)
        np.testing.assert_allclose(x.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-5, err_msg=name)
        # and against the reference's own fp32 numbers (golden), whose rounding error is ~3e-4 at nll~1e3
        np.testing.assert_allclose(nll.detach().cpu().numpy(), g[name + ".nll"], rtol=1e-4, atol=1e-3, err_msg=name)
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[name + ".grad"], rtol=1e-3, atol=1e-4, err_msg=name)


def test_ctc_module_reference_call_signature():
    """criterion(log_probs[T,B,C], targets, preds_size (CPU!), length) exactly as valid.py:33-38 / train.py:24-28."""
    h = _pkg()
    g = np.load(os.path.join(G, "ctc_cases.npz"))
    name = "iam_shape"
    logits, tg, tl = g[name + ".logits"], g[name + ".targets"], g[name + ".tgt_len"]
    B, T, C = logits.shape
    preds = torch.from_numpy(logits).cuda().requires_grad_(True)
    lp = preds.float().permute(1, 0, 2).log_softmax(2)
    crit = h.CTCLoss(reduction="none", zero_infinity=True).to("cuda")
    preds_size = torch.IntTensor([T] * B)                       # CPU tensor, as in valid.py
    loss = crit(lp, torch.from_numpy(tg).cuda(), preds_size, torch.from_numpy(tl).cuda()).mean()
    loss.backward()
    assert abs(loss.item() - g[name + ".nll"].mean()) <= 1e-4 * abs(g[name + ".nll"].mean())
    np.testing.assert_allclose(preds.grad.cpu().numpy() * B, g[name + ".grad"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("B,T,C,lo,hi", [(128, 128, 80, 16, 64), (32, 256, 90, 64, 200), (16, 128, 228, 1, 64)])
def test_ctc_full_size_properties(B, T, C, lo, hi):
    """Full BASELINE sizes: compare a subset against the float64 reference, and check size-independent
    properties on everything (rows of softmax - posterior sum to 0; nll finite and > 0)."""
    h = _pkg()
    rs = np.random.RandomState(B + T)
    logits = rs.randn(B, T, C).astype(np.float32)
    tl = rs.randint(lo, hi + 1, size=B).astype(np.int32)
    tg = rs.randint(1, C, size=int(tl.sum())).astype(np.int32)
    il = np.full(B, T, dtype=np.int32)
    x = torch.from_numpy(logits).cuda().requires_grad_(True)
    nll = h.ctc_loss_from_logits(x, torch.from_numpy(tg).cuda(), torch.from_numpy(tl), max_target_len=int(tl.max()))
    nll.sum().backward()
    gr = x.grad.cpu().numpy()
    assert np.isfinite(gr).all() and (nll > 0).all()
    assert np.abs(gr.sum(axis=2)).max() < 2e-4
    n = 8
    ref_nll, ref_grad = _ctc_ref64(logits[:n], tg[: int(tl[:n].sum())], il[:n], tl[:n])
    np.testing.assert_allclose(nll.detach().cpu().numpy()[:n], ref_nll, rtol=1e-4)
    np.testing.assert_allclose(gr[:n], ref_grad, rtol=1e-4, atol=1e-5)
    # worst-case provisioning path (no host knowledge of the label lengths): same numbers
    x2 = torch.from_numpy(logits).cuda().requires_grad_(True)
    nll2 = h.ctc_loss_from_logits(x2, torch.from_numpy(tg).cuda(), torch.from_numpy(tl).cuda())
    nll2.sum().backward()
    # (with the lane-group kernel forced, the unprovisioned call may be served by the other kernel: two
    # implementations, so equal to rounding rather than bit for bit)
    np.testing.assert_allclose(nll2.detach().cpu().numpy(), nll.detach().cpu().numpy(), rtol=2e-5)
    np.testing.assert_allclose(x2.grad.cpu().numpy(), gr, rtol=1e-4, atol=2e-6)


def _fallbacks():
    from importlib import import_module
    return import_module("htr-vt_b200._lib").lib().htrvt_ctc_fallback_count()


def _flagged():
    """sequences the lane-group kernel (fp32 linear domain) handed to the CTA-per-sequence kernel's fix-up launch"""
    from importlib import import_module
    return import_module("htr-vt_b200._lib").lib().htrvt_ctc_flagged_count()


def test_ctc_large_batch_kernels_agree():
    """The lane-group kernel (forced: the automatic switch sits at B >= 1024) against the CTA-per-sequence kernel
    on the same 1024 sequences, and against the float64 reference on a subset; infeasible / empty / ragged sequences mixed in."""
    from importlib import import_module
    h = _pkg()
    lib = import_module("htr-vt_b200._lib").lib()
    B, T, C = 1024, 128, 80
    rs = np.random.RandomState(7)
    logits = (rs.randn(B, T, C) * 1.5).astype(np.float32)
    tl = rs.randint(0, 65, size=B).astype(np.int32)
    il = rs.randint(40, T + 1, size=B).astype(np.int32)
    il[::3] = T
    tg = rs.randint(1, C, size=int(tl.sum())).astype(np.int32)
    tg[1::2] = tg[0::2][: len(tg[1::2])]                    # many adjacent repeats: some sequences become infeasible
    res = {}
    for mode in (0, 1):
        prev = lib.htrvt_ctc_set_mode(mode)
        x = torch.from_numpy(logits).cuda().requires_grad_(True)
        nll = h.ctc_loss_from_logits(x, torch.from_numpy(tg).cuda(), torch.from_numpy(tl),
                                     input_lengths=torch.from_numpy(il).cuda(), max_target_len=int(tl.max()))
        nll.sum().backward()
        res[mode] = (nll.detach().cpu().numpy(), x.grad.cpu().numpy())
        lib.htrvt_ctc_set_mode(prev)
    for mode in (1,):
        np.testing.assert_allclose(res[mode][0], res[0][0], rtol=2e-5, atol=1e-4)
        np.testing.assert_allclose(res[mode][1], res[0][1], rtol=1e-4, atol=2e-6)
    assert (res[1][0] == 0).sum() > 5 and (res[1][0] > 0).sum() > 500          # both kinds present
    sub = slice(0, 24)
    offs = np.concatenate([[0], np.cumsum(tl)])
    ref_nll, ref_grad = _ctc_ref64(logits[sub], tg[:offs[24]], il[sub], tl[sub])
    for mode in (1,):
        np.testing.assert_allclose(res[mode][0][sub], ref_nll, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(res[mode][1][sub], ref_grad, rtol=1e-4, atol=1e-5)


def test_ctc_fast_path_is_taken_and_fallback_is_exact():
    """The linear-domain fp64 recursion serves ordinary inputs; confidently-wrong logits (within-row spreads far
    beyond the fp64 range, which log space keeps) must be detected and recomputed in log space - same numbers."""
    h = _pkg()
    rs = np.random.RandomState(7)
    B, T, C = 16, 128, 80
    tl = rs.randint(20, 60, size=B).astype(np.int32)
    tg = rs.randint(1, C, size=int(tl.sum())).astype(np.int32)
    il = np.full(B, T, dtype=np.int32)
    # (a) ordinary and moderately peaked logits: no fallback
    for scale in (1.0, 8.0):
        logits = (rs.randn(B, T, C) * scale).astype(np.float32)
        n0, f0 = _fallbacks(), _flagged()
        x = torch.from_numpy(logits).cuda().requires_grad_(True)
        nll = h.ctc_loss_from_logits(x, torch.from_numpy(tg).cuda(), torch.from_numpy(tl))
        nll.sum().backward()
        torch.cuda.synchronize()
        assert _fallbacks() == n0, scale
        if scale == 1.0:                  # ordinary logits stay on the lane-group kernel's fp32 path too
            assert _flagged() == f0
        ref_nll, ref_grad = _ctc_ref64(logits, tg, il, tl)
        np.testing.assert_allclose(nll.detach().cpu().numpy(), ref_nll, rtol=1e-4)
        np.testing.assert_allclose(x.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-5)
    # (b) a "trained, confident" network: the labels' own alignment gets +25 logits (fast path, huge dynamic range
    # between on-path and off-path states, none of it relevant), then the same network on WRONG labels
    logits = rs.randn(B, T, C).astype(np.float32)
    pos = 0
    for b in range(B):
        lab = tg[pos:pos + tl[b]]
        pos += tl[b]
        frames = np.sort(rs.choice(T, size=tl[b], replace=False))
        logits[b, :, 0] += 25.0
        for f, c in zip(frames, lab):
            logits[b, f, 0] -= 25.0
            logits[b, f, c] += 25.0
    wrong = ((tg + rs.randint(1, C - 1, size=tg.shape)) % (C - 1) + 1).astype(np.int32)
    # (labels, logit scale, fallback expected?, grad atol).  x4 = +-100 logits on wrong labels: |log p| ~ 1e4, where
    # the log-space path itself (fp32, like ATen's) resolves ~3e-4; the fast path must hand such rows over.
    for labels, scale, expect_fallback, atol in ((tg, 1.0, False, 1e-5), (wrong, 1.0, None, 1e-5),
                                                 (wrong, 2.0, None, 2e-5), (wrong, 4.0, True, 5e-4)):
        lg = (logits * scale).astype(np.float32)
        n0, f0 = _fallbacks(), _flagged()
        x = torch.from_numpy(lg).cuda().requires_grad_(True)
        nll = h.ctc_loss_from_logits(x, torch.from_numpy(labels).cuda(), torch.from_numpy(tl))
        nll.sum().backward()
        torch.cuda.synchronize()
        if expect_fallback is not None:
            assert (_fallbacks() > n0) == expect_fallback, scale
        if labels is tg:                  # a confident network on its OWN labels: the fp32 lane-group path serves it
            assert _flagged() == f0
        ref_nll, ref_grad = _ctc_ref64(lg, labels, il, tl)
        np.testing.assert_allclose(nll.detach().cpu().numpy(), ref_nll, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(x.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=atol)


@pytest.mark.parametrize("T,C", [(128, 81), (37, 13), (5, 2), (256, 90)])
def test_ctc_odd_shapes_ragged_lengths(T, C):
    """Class counts that are not a multiple of 4 (scalar staging path), ragged input lengths (rows beyond a line's
    length get zero gradient), label lengths at the state-per-lane boundaries (S = 31, 33, 63, 65, 127, 129 ...) and
    lines that end while the posterior pass already runs (overlapped collection): nll / gradient vs float64."""
    h = _pkg()
    rs = np.random.RandomState(T * 7 + C)
    want_L = [0, 1, 15, 16, 31, 32, 47, 48, 63, 64, 65, 100]
    il = np.array([T, T, T, max(1, T - 1), T, max(1, T // 2 + 1), T, T, T, max(1, T - 3), T, T], dtype=np.int32)
    tl = np.array([min(L, int(t) // 2 if C == 2 else int(t) * 3 // 4) for L, t in zip(want_L, il)], dtype=np.int32)
    B = len(tl)
    labels = rs.randint(1, C, size=int(tl.sum())).astype(np.int32)
    lg = (rs.randn(B, T, C) * 2.0).astype(np.float32)
    x = torch.from_numpy(lg).cuda().requires_grad_(True)
    nll = h.ctc_loss_from_logits(x, torch.from_numpy(labels).cuda(), torch.from_numpy(tl), torch.from_numpy(il))
    nll.sum().backward()
    ref_nll, ref_grad = _ctc_ref64(lg, labels, il, tl)
    np.testing.assert_allclose(nll.detach().cpu().numpy(), ref_nll, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref_grad, rtol=1e-4, atol=2e-6)
    for b in range(B):
        assert not x.grad[b, int(il[b]):].any()


def test_decode_golden_strings():
    h = _pkg()
    g = np.load(os.path.join(G, "decode_cases.npz"))
    conv = h.CTCLabelConverter(str(g["alphabet"]))
    idx = torch.from_numpy(g["index"]).cuda()
    assert conv.decode(idx, torch.from_numpy(g["lens"])) == g["strings"].tolist()
    assert conv.decode(idx, torch.from_numpy(g["lens2"])) == g["strings2"].tolist()
    c87 = h.CTCLabelConverter("".join(chr(48 + i) for i in range(87)))
    t, l = c87.encode(["0a", "[x]"])
    assert t.cpu().tolist() == g["enc87_text"].tolist() and l.cpu().tolist() == g["enc87_len"].tolist()
    assert len(c87.character) == int(g["n_character87"])


def test_argmax_semantics_and_fused_decode():
    h = _pkg()
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    g = np.load(os.path.join(G, "argmax_cases.npz"))
    ids, lens, raw = ops.greedy_decode_ids(torch.from_numpy(g["logits"]).cuda(), 1000, want_raw=True)
    np.testing.assert_array_equal(raw.cpu().numpy(), g["index"])
    # fused decode == oracle collapse of oracle argmax, full size, bit exact
    rs = np.random.RandomState(3)
    B, T, C = 512, 128, 80
    logits = rs.randn(B, T, C).astype(np.float32)
    logits[:, :, 0] += 1.5                         # plenty of blanks and repeats
    logits = np.round(logits * 2) / 2              # many exact ties
    am = O.argmax_first(logits)
    want = O.greedy_ids(am.reshape(-1), [T] * B, 60)
    ids, lens = h.greedy_decode(torch.from_numpy(logits).cuda(), 60)
    ids, lens = ids.cpu().numpy(), lens.cpu().numpy()
    got = [ids[b, : lens[b]].tolist() for b in range(B)]
    assert got == want


@pytest.mark.gpu
@pytest.mark.parametrize("T,C", [(37, 7), (128, 80), (70, 13), (33, 4), (5, 1), (300, 90)])
def test_argmax_signed_zeros_nan_payloads_denormals(T, C):
    """The integer-key arg-max of the decode kernel against the oracle on what a key transform gets wrong first: -0 / +0
    ties, NaNs with the sign bit set, denormals, +-inf, rows shorter than a 16-byte group, ragged lengths that end inside
    a 32-row commit group."""
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    rs = np.random.RandomState(T * 131 + C)
    B = 9
    pool = np.array([0.0, -0.0, 1e-42, -1e-42, 1.0, -1.0, np.inf, -np.inf, 3.5, -3.5], dtype=np.float32)
    a = pool[rs.randint(0, len(pool), size=(B, T, C))]
    a[0] = np.where(rs.rand(T, C) < 0.5, np.float32(0.0), np.float32(-0.0))          # only signed zeros
    a[1] = -np.abs(a[1])                                                            # nothing above -0
    nanbits = np.array([0x7fc00000, 0xffc00000, 0xff800001, 0x7f800001], dtype=np.uint32).view(np.float32)
    m = rs.rand(T, C) < 0.1
    a[2][m] = nanbits[rs.randint(0, 4, size=int(m.sum()))]
    a[3] = -np.inf
    lengths = rs.randint(0, T + 1, size=B).astype(np.int32)
    lengths[0] = T
    want_raw = O.argmax_first(a)
    ids, lens, raw = ops.greedy_decode_ids(torch.from_numpy(a).cuda(), 1000, lengths=torch.from_numpy(lengths).cuda(),
                                           want_raw=True)
    raw, ids, lens = raw.cpu().numpy(), ids.cpu().numpy(), lens.cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(raw[b, : lengths[b]], want_raw[b, : lengths[b]])
    want = O.greedy_ids(np.concatenate([want_raw[b, : lengths[b]] for b in range(B)]), lengths.tolist(), 1000)
    assert [ids[b, : lens[b]].tolist() for b in range(B)] == want
    # a view with a row stride that is not a multiple of 4 floats takes the 4-byte staging path: same answer
    wide = torch.zeros(B, T, C + 1).cuda()
    wide[:, :, :C] = torch.from_numpy(a).cuda()
    _, _, raw2 = ops.greedy_decode_ids(wide[:, :, :C], 1000, want_raw=True)
    np.testing.assert_array_equal(raw2.cpu().numpy(), want_raw)


class _StubLM(object):
    """Same deterministic scorer as oracle/make_golden.py::StubLM (stands in for KenLM: any .score(text))."""

    def score(self, text):
        return sum(((ord(ch) * 31 + i * 17) % 97) / 97.0 for i, ch in enumerate(text)) - 0.6 * len(text)


def test_kbest_paths_match_oracle_and_reference_strings():
    """K-best per-frame beam (model_window/test_with_kenlm.py:25-59): candidate id rows and float64 scores against the
    oracle restatement, best strings against the fixtures produced by the reference's own function."""
    h = _pkg()
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    g = np.load(os.path.join(G, "beam_cases.npz"))
    alphabet = str(g["alphabet"])
    conv = h.CTCLabelConverter(alphabet)
    lm = _StubLM()
    for i in range(len(g["best"])):
        lp, K = g["log_probs"][i], int(g["beam"][i])
        x = torch.from_numpy(lp).cuda()
        assert h.simple_ctc_beam_search_with_lm(x, conv, lm, beam_size=K) == str(g["best"][i])
        ids, lens, sc = ops.ctc_kbest_paths(x.unsqueeze(1), K)
        want = O.kbest_paths(lp, K)
        got = [(ids[0, r, : int(lens[0, r])].cpu().tolist(), float(sc[0, r])) for r in range(K)]
        assert [a[0] for a in got] == [w[0] for w in want]
        np.testing.assert_allclose([a[1] for a in got], [w[1] for w in want], rtol=0, atol=1e-9)
    # one launch for a batch in the reference's [T, B, C] layout == line by line
    lpb = torch.from_numpy(np.ascontiguousarray(g["log_probs"][0::3].transpose(1, 0, 2))).cuda()   # the K = 5 cases
    assert h.beam_search_with_lm_batch(lpb, conv, lm, beam_size=5) == [str(s) for s in g["best"][0::3]]


@pytest.mark.parametrize("T,C,K", [(1, 7, 5), (9, 3, 8), (128, 80, 5), (256, 90, 8), (64, 228, 1)])
def test_kbest_paths_edge_shapes(T, C, K):
    """T = 1, fewer classes than beams (C < K), the BASELINE line shapes, K = 1 (= the greedy path), ragged lengths,
    and exact ties (quantised log-probs): ids, lens, scores and the order of equal-score beams match the oracle."""
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    rs = np.random.RandomState(T + C + K)
    B = 5
    x = rs.randn(B, T, C).astype(np.float32) * 2.0
    x[1] = np.round(x[1] * 2) / 2                                   # many exactly equal candidates
    x[2, :, 0] += 5.0
    lengths = np.array([T, T, max(1, T // 2), T, max(1, T - 1)], dtype=np.int32)
    ids, lens, sc = ops.ctc_kbest_paths(torch.from_numpy(x).cuda(), K, torch.from_numpy(lengths), layout="btc")
    ids, lens, sc = ids.cpu().numpy(), lens.cpu().numpy(), sc.cpu().numpy()
    for b in range(B):
        want = O.kbest_paths(x[b, : lengths[b]], K)
        assert (lens[b] >= 0).sum() == len(want)
        for r, (text, score) in enumerate(want):
            assert ids[b, r, : lens[b, r]].tolist() == text, (b, r)
            assert abs(sc[b, r] - score) < 1e-9
            assert not ids[b, r, lens[b, r]:].any()
    if K == 1:                                                       # the single best path is the arg-max path
        am = O.argmax_first(x)
        for b in range(B):
            want = [int(v) for i, v in enumerate(am[b, : lengths[b]]) if v != 0 and not (i > 0 and am[b, i - 1] == v)]
            if not (np.round(x[b] * 2) / 2 == x[b]).all():          # (ties: argmax takes the LOWEST index, argsort the highest)
                assert ids[b, 0, : lens[b, 0]].tolist() == want


@pytest.mark.parametrize("T,C,K", [(1, 7, 5), (9, 3, 8), (6, 4, 16), (128, 80, 5), (128, 80, 16), (256, 90, 8),
                                   (64, 228, 1), (40, 33, 3)])
def test_prefix_beam_matches_oracle(T, C, K):
    """CTC prefix beam search kernel (csrc/prefix_beam.cu) vs the float64 oracle restatement (itself pinned to a
    brute-force enumeration, tests/test_oracle.py): labellings, their order and float64 scores for every surviving
    entry, on random / peaked / blank-heavy / quantised (exact ties) log-probs and ragged lengths."""
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    rs = np.random.RandomState(7 * T + C + K)
    B = 6
    x = rs.randn(B, T, C).astype(np.float32) * 2.0
    x[1] = np.round(x[1] * 2) / 2                                   # many exactly equal candidates
    x[2, :, 0] += 5.0                                               # blank-dominated
    x[3] *= 4.0                                                     # peaked: merges of doubled labels matter
    x[5] = np.repeat(x[5, ::2], 2, axis=0)[:T]                      # every frame twice: repeats / stay-vs-extend
    lp = torch.from_numpy(x).log_softmax(-1).numpy()
    lengths = np.array([T, T, max(1, T // 2), T, max(1, T - 1), T], dtype=np.int32)
    ids, lens, sc = ops.ctc_prefix_beam(torch.from_numpy(lp).cuda(), K, torch.from_numpy(lengths), layout="btc")
    ids, lens, sc = ids.cpu().numpy(), lens.cpu().numpy(), sc.cpu().numpy()
    for b in range(B):
        want = O.ctc_prefix_beam_search(lp[b, : lengths[b]], K)
        assert (lens[b] >= 0).sum() == len(want), b
        for r, (lab, score) in enumerate(want):
            assert ids[b, r, : lens[b, r]].tolist() == lab, (b, r)
            if np.isfinite(score):
                assert abs(sc[b, r] - score) < 1e-9 * max(1.0, abs(score)), (b, r)
            else:
                assert sc[b, r] == score
            assert not ids[b, r, lens[b, r]:].any()
        assert (lens[b, len(want):] == -1).all() and not ids[b, len(want):].any()
    # [T, B, C] layout without lengths == [B, T, C]
    ids2, lens2, sc2 = ops.ctc_prefix_beam(torch.from_numpy(np.ascontiguousarray(lp.transpose(1, 0, 2))).cuda(), K)
    full = [b for b in range(B) if lengths[b] == T]
    assert (ids2.cpu().numpy()[full] == ids[full]).all() and (lens2.cpu().numpy()[full] == lens[full]).all()


def test_prefix_beam_exact_posterior_and_lm_entry_points():
    """Small lines where K = 16 holds every prefix: the kernel's scores are the brute-force labelling posterior.
    The LM entry points with search='prefix' agree with the oracle's pick; doubled letters survive."""
    h = _pkg()
    from importlib import import_module
    ops = import_module("htr-vt_b200.ops")
    rs = np.random.RandomState(11)
    for T, C in [(2, 3), (3, 3), (5, 2)]:
        x = rs.randn(1, T, C).astype(np.float32) * 1.5
        lp = torch.from_numpy(x).log_softmax(-1)
        exact = O.labelling_logprobs_bruteforce(lp[0].numpy())
        ids, lens, sc = ops.ctc_prefix_beam(lp.cuda(), 16, layout="btc")
        got = {tuple(ids[0, r, : int(lens[0, r])].cpu().tolist()): float(sc[0, r]) for r in range(16) if lens[0, r] >= 0}
        for lab, s in exact.items():
            assert abs(got[lab] - s) < 1e-9, (T, C, lab)
    alphabet = "abcdefghij"
    conv = h.CTCLabelConverter(alphabet)
    lm = _StubLM()
    x = rs.randn(40, 7, len(alphabet) + 1).astype(np.float32) * 3
    lp = torch.from_numpy(x).log_softmax(-1)
    got = h.beam_search_with_lm_batch(lp.cuda(), conv, lm, beam_size=8, search="prefix")
    want = [O.prefix_beam_search_with_lm(lp[:, b].numpy(), alphabet, lm.score, 8) for b in range(7)]
    assert got == want
    assert h.simple_ctc_beam_search_with_lm(lp[:, 2].cuda(), conv, lm, beam_size=8, search="prefix") == want[2]
    peaked = np.full((7, 4), -30.0, dtype=np.float32)
    for t, c in enumerate([1, 1, 0, 1, 2, 2, 0]):
        peaked[t, c] = 0.0
    cands = h.kbest_candidates(torch.from_numpy(peaked).unsqueeze(1).cuda(), h.CTCLabelConverter("abc"), 3,
                               search="prefix")[0]
    assert cands[0][0] == "aab"


def test_decode_logits_async_matches_sync_and_reuses_pinned_buffers():
    """CTCLabelConverter.decode_logits_async: same strings as decode_logits, several handles in flight, the pinned
    host pair of a resolved handle is reused by the next call."""
    import htrvt_b200 as h
    conv = h.CTCLabelConverter("".join(chr(33 + i) for i in range(79)))
    torch.manual_seed(2)
    batches = [torch.randn(64, 128, 80, device="cuda") * 3 for _ in range(4)]
    want = [conv.decode_logits(x) for x in batches]
    handles = [conv.decode_logits_async(x) for x in batches[:3]]            # three in flight
    got = [hd.strings() for hd in handles]
    assert got == want[:3]
    assert handles[0].strings() is got[0] and handles[0].done()             # resolved once, cached
    pool = conv._pinned[((64, 128), torch.cuda.current_device())]
    assert len(pool) == 3
    hd = conv.decode_logits_async(batches[3])
    assert len(pool) == 2                                                   # took a pinned pair back out of the pool
    assert hd.strings() == want[3] and len(pool) == 3
    lens = torch.randint(1, 129, (64,), dtype=torch.int32)
    assert conv.decode_logits_async(batches[0], lens).strings() == conv.decode_logits(batches[0], lens)
