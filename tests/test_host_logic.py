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
