"""GPU parity: tcgen05 attention fwd/bwd vs an fp32 torch reference of Attention.forward
(model_v1/model/HTR_VT.py:32-36) on the same bf16 q, k, v."""
from importlib import import_module

import pytest
import torch

pytestmark = pytest.mark.gpu


def ops():
    import htrvt_b200  # noqa: F401
    return import_module("htr-vt_b200.ops")


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize("B,H,T", [(2, 6, 128), (128, 6, 128), (3, 2, 32), (2, 6, 96)])
def test_attention_fwd_bwd(B, H, T):
    o = ops()
    hd = 128
    torch.manual_seed(0)
    qkv_tm = (torch.randn(B, T, 3, H, hd, device="cuda") * 0.7).bfloat16()     # token-major, as the QKV GEMM writes it
    qkv = qkv_tm.permute(2, 0, 3, 1, 4)
    scale = hd ** -0.5
    out = torch.empty(B, T, H * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda")
    o.attention_fwd(qkv_tm, out, lse, scale)
    q, k, v = [t.float().requires_grad_(True) for t in qkv]
    s = (q @ k.transpose(-2, -1)) * scale
    p = s.softmax(-1)
    ref = (p @ v).transpose(1, 2).reshape(B, T, H * hd)
    assert _rel(out, ref) < 1.5e-2
    assert _rel(lse, torch.logsumexp(s, -1)) < 1e-4
    dout = torch.randn(B, T, H * hd, device="cuda").bfloat16()
    ref.backward(dout.float())
    dqkv = torch.empty(B, T, 3, H, hd, device="cuda", dtype=torch.bfloat16)
    o.attention_bwd(qkv_tm, out, dout, lse, dqkv, scale)
    for i, t in enumerate((q, k, v)):
        want = t.grad.permute(0, 2, 1, 3)          # [B,T,H,hd]
        assert _rel(dqkv[:, :, i], want) < 2e-2, i
