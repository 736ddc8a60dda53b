"""fp32-parity mode (csrc/exact.cu + split-bf16 tensor-core contractions): north_star's fp32 bounds.
  * logits within 1e-4 (max-abs / max-abs) of the UNMODIFIED reference's fp32 logits (committed goldens);
  * greedy decode strings identical to the strings the reference's valid.py:40-42 + CTCLabelConverter.decode produced;
  * the building blocks against torch fp32 / float64."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import htrvt_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def ops():
    import htrvt_b200  # noqa: F401
    return import_module("htr-vt_b200.ops")


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_split3_is_exact_to_24_bits():
    o = ops()
    torch.manual_seed(0)
    x = torch.randn(3, 1000, device="cuda") * torch.logspace(-6, 6, 1000, device="cuda")
    p = o.split3(x)
    assert p.shape == (3, 3, 1000) and p.dtype == torch.bfloat16
    back = p[0].double() + p[1].double() + p[2].double()
    assert float(((back - x.double()).abs() / x.double().abs()).max()) < 2.0 ** -23


@pytest.mark.parametrize("M,N,K", [(256, 2304, 768), (384, 768, 3072), (130, 80, 768)])
def test_split_gemm_matches_float64(M, N, K):
    o = ops()
    torch.manual_seed(1)
    x = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    out = torch.full((M, N), 3.0, device="cuda")
    o.gemm_tn_split(o.split3(x), o.split3(w), out, b)
    ref = x.double() @ w.double().t() + b.double()
    # (the tensor pipe's fp32 accumulator truncates: ~1e-6 per 1000 terms of K, tools/acc_probe.py)
    assert _rel(out, ref) < 1e-5, _rel(out, ref)
    # the plain bf16 product of the same operands, for scale: four orders of magnitude coarser
    out16 = torch.empty(M, N, device="cuda")
    o.gemm_tn(x.bfloat16(), w.bfloat16(), out16, bias=b)
    assert _rel(out16, ref) > 1e-4


@pytest.mark.parametrize("NB,H,W,Cin,Cout,ks,sh,sw", [(2, 8, 256, 192, 192, 3, 1, 1), (2, 8, 256, 192, 384, 3, 2, 2),
                                                      (2, 8, 256, 192, 384, 1, 2, 2), (1, 2, 128, 768, 768, 3, 1, 1)])
def test_split_conv_matches_fp32(NB, H, W, Cin, Cout, ks, sh, sw):
    o = ops()
    torch.manual_seed(2)
    x = torch.relu(torch.randn(NB, H, W, Cin, device="cuda"))
    w = torch.randn(Cout, Cin, ks, ks, device="cuda") / (Cin * ks * ks) ** 0.5
    y = o.conv_fwd_split(o.split3(x), o.split3(w.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin)), ks, sh, sw)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, (sh, sw), ks // 2).permute(0, 2, 3, 1)
    assert y.dtype == torch.float32 and y.shape == ref.shape
    assert _rel(y, ref) < 1e-5, _rel(y, ref)


def test_fp32_elementwise_steps():
    o = ops()
    torch.manual_seed(3)
    # BatchNorm(eval) + residual / second BN + ReLU
    P, C = 500, 64
    raw, raw2, res = (torch.randn(P, C, device="cuda") for _ in range(3))
    st = torch.randn(4, C, device="cuda")
    st2 = torch.randn(4, C, device="cuda")
    y, pl = o.bn_act_f32(raw, st, True, res=res)
    ref = torch.relu(raw * st[2] + st[3] + res)
    assert _rel(y, ref) < 1e-6 and _rel(pl.double().sum(0), ref) < 1e-6
    y, _ = o.bn_act_f32(raw, st, True, raw2=raw2, st2=st2, want_planes=False)
    assert _rel(y, torch.relu(raw * st[2] + st[3] + raw2 * st2[2] + st2[3])) < 1e-6
    # final max-pool
    x = torch.randn(3, 2, 40, 32, device="cuda")
    assert torch.equal(o.maxpool_f32(x), F.max_pool2d(x.permute(0, 3, 1, 2), 3, (2, 1), 1).permute(0, 2, 3, 1))
    # tokens
    B, T, D = 2, 16, 32
    tok, mt, pos = torch.randn(B, T, D, device="cuda"), torch.randn(D, device="cuda"), torch.randn(T, D, device="cuda")
    mask = (torch.rand(T, device="cuda") > 0.4).float()
    m = mask.view(1, T, 1)
    assert _rel(o.tokens_f32(tok, mask, mt, pos, B, T, D).view(B, T, D), tok * m + (1 - m) * mt + pos) < 1e-6
    # row LayerNorm with fused residual add
    M, D = 100, 768
    x, ad, g, b = (torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda"),
                   torch.randn(D, device="cuda"), torch.randn(D, device="cuda"))
    pl, xn = o.row_ln_f32(x, g, b, 1e-6, ad)
    assert torch.equal(xn, x + ad)
    assert _rel(pl.double().sum(0), F.layer_norm((x + ad).double(), (D,), g.double(), b.double(), 1e-6)) < 2e-6
    # erf GELU
    u = torch.randn(1000, 64, device="cuda") * 3
    assert _rel(o.gelu_split(u).double().sum(0), F.gelu(u.double())) < 1e-6
    # attention
    B, H, T, hd = 2, 3, 128, 128
    qkv = torch.randn(B, T, 3, H, hd, device="cuda")
    out = o.attention_f32(qkv, B, H, T, hd, hd ** -0.5).view(B, T, H, hd)
    q, k, v = (qkv[:, :, i].double().permute(0, 2, 1, 3) for i in range(3))
    ref = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    assert _rel(out.permute(0, 2, 1, 3), ref) < 2e-6


def _build_full(seed):
    H = import_module("htr-vt_b200.model.HTR_VT")
    m = H.create_model(80, [64, 512])
    sd = O.init_state_dict(80, [64, 512], seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def test_fp32_mode_logits_within_1e4_of_reference():
    import htrvt_b200  # noqa: F401
    g = np.load(os.path.join(G, "v1_full.npz"))
    nb_cls, W, B, seed = [int(v) for v in g["meta"][:4]]
    m, _ = _build_full(seed)
    m.set_precision("fp32")
    x = torch.from_numpy(np.random.RandomState(seed + 1).rand(B, 1, 64, W).astype(np.float32)).cuda()
    with torch.no_grad():
        got = m(x)
    assert got.dtype == torch.float32 and got.shape == (B, W // 4, nb_cls)
    err = _rel(got.cpu(), torch.from_numpy(g["logits_eval"]))
    assert err < 1e-4, err
    with pytest.raises(Exception):                       # training / autograd need the 16-bit path
        m.train()
        m(x)
    m.eval().set_precision("bf16")
    with torch.no_grad():
        assert _rel(m(x).float().cpu(), torch.from_numpy(g["logits_eval"])) < 2e-2


def test_fp32_mode_decode_strings_equal_the_reference_end_to_end():
    """images -> our encoder (fp32-parity mode) -> our greedy decode == the strings the reference
    model + valid.py:40-42 + CTCLabelConverter.decode produced on the same 32 lines (golden v1_train_b32.strings_eval)."""
    import htrvt_b200 as h
    g = np.load(os.path.join(G, "v1_train_b32.npz"))
    nb_cls, W, B, seed, _ = [int(v) for v in g["meta"]]
    m, _ = _build_full(seed)
    m.set_precision("fp32")
    x = torch.from_numpy(np.random.RandomState(seed + 1).rand(B, 1, 64, W).astype(np.float32)).cuda()
    alphabet = "".join(chr(33 + i) for i in range(nb_cls - 1))
    conv = h.CTCLabelConverter(alphabet)
    with torch.no_grad():
        preds = m(x).float()
        assert _rel(preds.cpu(), torch.from_numpy(g["logits_eval"])) < 1e-4
        # the reference's own call sequence (valid.py:32-42) on the drop-in objects ...
        lp = preds.permute(1, 0, 2).log_softmax(2)
        _, idx = lp.max(2)
        idx = idx.transpose(1, 0).contiguous().view(-1)
        strings = conv.decode(idx.data, torch.IntTensor([preds.size(1)] * B))
        # ... and the fused device decode
        fused = conv.decode_logits(preds)
    want = g["strings_eval"].tolist()
    assert strings == want
    assert fused == want
    assert np.array_equal(idx.cpu().numpy(), g["index_eval"].astype(np.int64))
