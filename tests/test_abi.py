"""CPU-only: the C-ABI library loads and exports every symbol include/htrvt.h declares (no compute calls),
the python boundary refuses to run without CUDA, and host-side logic (label encoding, DP sharding) is right."""
import ctypes
import os
import re
from importlib import import_module

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import htrvt_b200  # noqa: F401
    return import_module("htr-vt_b200._lib")


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "htrvt.h")).read()
    names = sorted(set(re.findall(r"\b(htrvt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    L = _lib()
    lib = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the python side declares a signature for everything the header exports (except the pure queries)
    undeclared = [n for n in names if n not in L.SIGNATURES and n != "htrvt_launch_count" and n != "htrvt_conv1_fwd_zdim"]
    assert not undeclared, undeclared
    assert lib.htrvt_version() >= 100


def test_no_cpu_fallback():
    import htrvt_b200 as h
    H = import_module("htr-vt_b200.model.HTR_VT")
    m = H.create_model(80, [64, 512])
    with pytest.raises(h._lib.HtrvtError):
        m(torch.zeros(1, 1, 64, 512))
    with pytest.raises(h._lib.HtrvtError):
        h.ctc_loss_from_logits(torch.zeros(1, 4, 5), torch.tensor([1]), torch.tensor([1]))
    with pytest.raises(h._lib.HtrvtError):
        h.greedy_decode(torch.zeros(1, 4, 5), 5)


def test_state_dict_schema_and_init():
    H = import_module("htr-vt_b200.model.HTR_VT")
    import htrvt_oracle as O
    m = H.create_model(80, [64, 512])
    assert float(m.head.bias.abs().max()) == 0.0                        # reference zero-inits Linear biases
    sd = O.init_state_dict(80, [64, 512], seed=3)
    assert list(m.state_dict().keys()) == [k for k in m.state_dict().keys()]
    assert sorted(m.state_dict().keys()) == sorted(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    np.testing.assert_allclose(m.pos_embed.numpy()[0], O.sincos_pos_embed(768, [16, 8]), atol=1e-6)
    assert not m.pos_embed.requires_grad
    assert sum(p.numel() for p in m.parameters()) == 53486096          # SURVEY.md 8a [probe]
    g = np.load(os.path.join(ROOT, "tests", "golden", "v1_full.npz"))
    assert list(m.state_dict().keys()) == g["keys"].tolist()            # reference key ORDER too


def test_span_mask_draw_parity():
    H = import_module("htr-vt_b200.model.HTR_VT")
    import htrvt_oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "v1_full.npz"))
    torch.manual_seed(7)
    a = H.MaskedAutoencoderViT.span_mask(128, 0.4, 8)
    np.testing.assert_array_equal(a.numpy(), g["mask"])                 # the reference's own draw (seed 7)
    torch.manual_seed(99)
    a = H.MaskedAutoencoderViT.span_mask(256, 0.4, 8)
    torch.manual_seed(99)
    np.testing.assert_array_equal(a.numpy(), O.draw_span_mask(256, 0.4, 8).numpy())


def test_converter_encode_matches_reference_golden():
    import htrvt_b200 as h
    g = np.load(os.path.join(ROOT, "tests", "golden", "decode_cases.npz"))
    c87 = h.CTCLabelConverter("".join(chr(48 + i) for i in range(87)), device="cpu")
    t, l = c87.encode(["0a", "[x]"])
    assert t.tolist() == g["enc87_text"].tolist() and l.tolist() == g["enc87_len"].tolist()
    assert len(c87.character) == int(g["n_character87"])


def test_dp_helpers():
    ddp = import_module("htr-vt_b200.ddp")
    assert [ddp.shard_batch(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    assert ddp.shard_batch(10, 3, 4) == (9, 10) and ddp.shard_batch(2, 3, 4) == (2, 2)
    names = ["mask_token", "pos_embed", "patch_embed.conv1.weight", "blocks.0.norm1.weight", "head.bias"]
    offs, split = ddp.segment_bounds(names, [768, 10, 1728, 768, 80])
    # every tensor starts on a 64-float (256 B) boundary: TMA reduce-add targets need 16-byte aligned bases
    assert offs == [0, 768, 832, 2560, 3328, 3456] and split == 2560


def test_reference_checkpoint_round_trip(tmp_path):
    """The reference's checkpoint layout (train.py:153-172) and its loading recipe (test.py:29-40: 'state_dict_ema',
    `module.` prefixes, strict=True) work on the drop-in module; CPU only (no kernels involved)."""
    import copy
    from importlib import import_module
    import torch
    H = import_module("htr-vt_b200.model.HTR_VT")
    C = import_module("htr-vt_b200.utils.checkpoint")
    torch.manual_seed(0)
    m = H.create_model(20, [64, 128])
    ema = copy.deepcopy(m)
    with torch.no_grad():
        for p in ema.parameters():
            p.add_(1.0)
    path = str(tmp_path / "best_CER.pth")
    wrapped = {"module." + k: v for k, v in ema.state_dict().items()}          # as saved from a DataParallel wrapper
    torch.save({"model": m.state_dict(), "state_dict_ema": wrapped, "nb_iter": 7, "best_cer": 0.05}, path)
    m2 = H.create_model(20, [64, 128])
    ckpt = C.load_reference_checkpoint(m2, path)
    assert ckpt["nb_iter"] == 7
    for (k, a), (_, b) in zip(m2.state_dict().items(), ema.state_dict().items()):
        assert torch.equal(a, b), k
    C.load_reference_checkpoint(m2, path, prefer_ema=False)
    for (k, a), (_, b) in zip(m2.state_dict().items(), m.state_dict().items()):
        assert torch.equal(a, b), k
    out = C.save_reference_checkpoint(str(tmp_path / "x.pth"), m, model_ema=ema, nb_iter=3)
    assert list(out["state_dict_ema"].keys()) == list(m.state_dict().keys()) and out["nb_iter"] == 3
