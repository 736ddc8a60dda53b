"""Host-side logic that needs no GPU: candidate string building of the LM-rescoring decode, converter encode,
batch sharding.  (The device kernels are covered by the -m gpu tests; nothing here calls into the CUDA library.)"""
from importlib import import_module

import numpy as np
import torch

import htrvt_oracle as O


def test_beam_candidate_strings_match_reference_decode():
    """htr-vt_b200/beam.py::_candidate_strings == the reference converter's decode (model_v1/utils/utils.py:72-86,
    restated in the oracle) applied to every already-collapsed candidate row: blanks / repeats / out-of-alphabet ids
    are filtered AGAIN, rows are cut at their length, empty rows give ''."""
    import htrvt_b200  # noqa: F401
    beam = import_module("htr-vt_b200.beam")

    class Conv(object):
        character = ["[blank]"] + list("abcde")

    rs = np.random.RandomState(0)
    ids = rs.randint(0, 9, size=(64, 17)).astype(np.int32)          # ids 6..8 are beyond the alphabet
    lens = rs.randint(0, 18, size=64)
    got = beam._candidate_strings(ids, lens, Conv())
    want = [O.decode_strings(ids[i, :lens[i]].astype(np.int64), [int(lens[i])], "abcde")[0] for i in range(64)]
    assert got == want
    assert beam._pick([], None) == ""

    class LM(object):
        def score(self, text):
            return -abs(len(text) - 3)

    assert beam._pick([("a", -1.0), ("abc", -9.0), ("abcd", -0.5)], LM()) == "abc"


def test_converter_encode_matches_reference_layout():
    import htrvt_b200 as h
    conv = h.CTCLabelConverter("abc", device="cpu")
    text, length = conv.encode(["ab", "", "cab"])
    assert text.tolist() == [1, 2, 3, 1, 2] and length.tolist() == [2, 0, 3]
    assert text.dtype == torch.int32 and length.dtype == torch.int32
    assert conv.character[0] == "[blank]" and len(conv.character) == 4


def test_shard_batch_covers_every_item_once():
    ddp = import_module("htr-vt_b200.ddp")
    for n, world in [(4096, 8), (10, 4), (3, 8), (0, 2)]:
        seen = []
        for r in range(world):
            lo, hi = ddp.shard_batch(n, r, world)
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_ema_update_on_cpu_state_dicts_follows_the_reference_expression():
    """ema_update_ on tensors the fused kernel does not take (CPU fp32, int64 counters, fp64): every entry goes through
    model_v1/utils/utils.py:173's expression `ema_v.copy_(ema_v * decay + (1 - decay) * model_v)`, counters included
    (float math, truncating copy back)."""
    import copy
    U = import_module("htr-vt_b200.utils.utils")
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 4, 3),
                              torch.nn.BatchNorm2d(4))
    net.register_buffer("wide", torch.randn(5, dtype=torch.float64))
    shadow = copy.deepcopy(net)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(torch.randn_like(p))
        net[1].num_batches_tracked.add_(7)
        net[3].num_batches_tracked.add_(123456789)
        net[1].running_var.mul_(3.0)
        net.wide.add_(1.0)
    want = {k: v.clone() for k, v in shadow.state_dict().items()}
    for d in (0.1, 0.9999):
        src = net.state_dict()
        for k in want:
            want[k] = (want[k] * d + (1. - d) * src[k]).to(want[k].dtype)
        U.ema_update_(shadow, net, d)
    got = shadow.state_dict()
    assert list(got) == list(want)
    for k in want:
        assert got[k].dtype == want[k].dtype
        assert torch.equal(got[k], want[k]), k
    assert U.effective_decay(0.9999, 0) == 0.1 and U.effective_decay(0.9999, -1) == 0.9999


def test_sam_refuses_cpu_parameters_loudly():
    """No CPU fallback on the product path: the multi-tensor SAM raises on parameters the kernels cannot take instead of
    quietly running another implementation; parameters without a gradient are skipped like the reference does
    (model_v1/utils/sam.py:20)."""
    import pytest
    S = import_module("htr-vt_b200.utils.sam")
    w = torch.nn.Parameter(torch.randn(4, 3))
    frozen = torch.nn.Parameter(torch.randn(2))
    opt = S.SAM([w, frozen], torch.optim.AdamW, lr=1e-3)
    ps, gs = opt._live(opt.param_groups[0])
    assert ps == [] and gs == []                       # nothing has a gradient yet
    w.grad = torch.randn_like(w)
    with pytest.raises(RuntimeError, match="contiguous fp32 CUDA"):
        opt.first_step()
    with pytest.raises(RuntimeError, match="contiguous fp32 CUDA"):
        opt.second_step()
