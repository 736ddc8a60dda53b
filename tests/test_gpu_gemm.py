"""GPU parity: the tcgen05 tap-GEMM (linear fwd/dgrad/wgrad, conv fwd/dgrad/wgrad) through the C ABI
against fp32 torch references computed from the same bf16 operands."""
from importlib import import_module

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ops():
    import htrvt_b200  # noqa: F401
    return import_module("htr-vt_b200.ops")


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (16384, 2304, 768), (16384, 768, 3072), (1000, 80, 768),
                                   (16384, 3072, 768), (384, 192, 192)])
def test_gemm_tn_bias(M, N, K):
    o = ops()
    torch.manual_seed(0)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    ref = x.float() @ w.float().t() + b
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o.gemm_tn(x, w, out, bias=b)
    assert _rel(out, ref) < 1e-2
    out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    o.gemm_tn(x, w, out32, bias=b)
    assert _rel(out32, ref) < 1e-4


@pytest.mark.parametrize("M,N,K", [(16384, 3072, 768), (384, 1024, 256), (130, 256, 64), (256, 768, 256)])
def test_gemm_fused_gelu_forward_and_backward(M, N, K):
    """timm Mlp's fc1 -> nn.GELU (erf form) fused into the GEMM epilogue (model_v1/model/HTR_VT.py:76): single store
    (eval), dual store with the pre-activation (train), and the activation's backward fused into fc2's input-gradient
    GEMM: dX = (dY W) * gelu'(u)."""
    o = ops()
    torch.manual_seed(5)
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda") * 0.5
    u_ref = x.float() @ w.float().t() + b
    a_ref = F.gelu(u_ref)
    a1 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    o.gemm_tn(x, w, a1, bias=b, gelu=True)                                  # eval: the activation only
    assert _rel(a1, a_ref) < 1e-2
    if N % 256 == 0:
        a2 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
        u2 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
        o.gemm_tn(x, w, a2, bias=b, gelu=True, pre=u2)                      # train: activation + its derivative
        assert _rel(a2, a1) < 4e-3                                          # (same formula, one shared evaluation)
        ur = u_ref.clone().requires_grad_(True)
        F.gelu(ur).sum().backward()
        assert _rel(u2, ur.grad) < 1e-2
    # backward: N here plays fc1's hidden width; dY [M, Kd] @ W2 [Kd, N]
    Kd = 256
    dy = torch.randn(M, Kd, device="cuda").bfloat16()
    w2 = (torch.randn(Kd, N, device="cuda") / Kd ** 0.5).bfloat16()
    u = (torch.randn(M, N, device="cuda") * 1.5).bfloat16()
    uf = u.float().requires_grad_(True)
    F.gelu(uf).backward(dy.float() @ w2.float())
    du = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    bias_grad = torch.ones(N, device="cuda")
    ud = uf.detach().clone().requires_grad_(True)
    F.gelu(ud).sum().backward()
    gd = ud.grad.bfloat16()                                                 # what the forward epilogue saves: gelu'(u)
    o.gemm_nn(dy, w2, du, gelu_u=gd, colsum=bias_grad)                      # + fc1's bias gradient from the epilogue
    assert _rel(du, uf.grad) < 1.5e-2
    assert _rel(bias_grad - 1.0, du.float().sum(0)) < 1e-4
    da = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o.gemm_nn(dy, w2, da)
    assert _rel(o.gelu_bwd(da, u), uf.grad) < 1.5e-2                         # the stand-alone pair agrees
    assert _rel(o.mul_bf16(da, gd), uf.grad) < 1.5e-2                        # ... and the multiply by the saved factor


def test_gemm_tn_relu_accumulate():
    o = ops()
    torch.manual_seed(1)
    M, N, K = 1024, 3072, 768
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda") * 0.1
    pre = x.float() @ w.float().t() + b
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o.gemm_tn(x, w, out, bias=b, relu=True)
    assert _rel(out, torch.relu(pre)) < 1e-2
    acc = torch.randn(M, N, device="cuda")
    base = acc.clone()
    o.gemm_tn(x, w, acc, bias=b, accumulate=True)           # TMA reduce-add into fp32
    assert _rel(acc, base + pre) < 1e-4
    accb = torch.randn(M, N, device="cuda").bfloat16()
    baseb = accb.clone()
    o.gemm_tn(x, w, accb, accumulate=True)                  # TMA reduce-add into bf16
    assert _rel(accb, baseb.float() + (pre - b)) < 1.5e-2
    # strided (column-sliced) output view
    big = torch.zeros(M, N + 64, device="cuda", dtype=torch.bfloat16)
    o.gemm_tn(x, w, big[:, :N], bias=b)
    assert _rel(big[:, :N], pre) < 1e-2 and float(big[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K", [(16384, 768, 2304), (1024, 768, 80), (16384, 3072, 768), (300, 192, 384)])
def test_gemm_nn_dgrad(M, N, K):
    o = ops()
    torch.manual_seed(2)
    Kp = (K + 7) // 8 * 8
    dy = torch.zeros(M, Kp, device="cuda", dtype=torch.bfloat16)
    dy[:, :K] = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(K, N, device="cuda") / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    o.gemm_nn(dy[:, :K], w, out)
    ref = dy[:, :K].float() @ w.float()
    assert _rel(out, ref) < 1e-2
    out32 = torch.randn(M, N, device="cuda")
    base = out32.clone()
    o.gemm_nn(dy[:, :K], w, out32, accumulate=True)
    assert _rel(out32, ref + base) < 1e-4


@pytest.mark.parametrize("M,N,K", [(16384, 2304, 768), (16384, 768, 3072), (4096, 80, 768), (1000, 192, 192)])
def test_linear_wgrad(M, N, K):
    o = ops()
    torch.manual_seed(3)
    Np = (N + 7) // 8 * 8
    dy = torch.randn(M, Np, device="cuda").bfloat16()
    x = torch.randn(M, K, device="cuda").bfloat16()
    g = torch.randn(N, K, device="cuda")
    base = g.clone()
    o.linear_wgrad(dy[:, :N], x, g, accumulate=True)
    ref = dy[:, :N].float().t() @ x.float()
    assert _rel(g - base, ref) < 2e-4
    o.linear_wgrad(dy[:, :N], x, g, accumulate=False)
    assert _rel(g, ref) < 2e-4


CONVS = [  # NB, H, W, Cin, Cout, ks, sh, sw
    (2, 16, 512, 192, 192, 3, 2, 1),
    (2, 8, 512, 192, 192, 3, 1, 1),
    (2, 8, 512, 192, 384, 3, 2, 2),
    (2, 4, 256, 384, 384, 3, 1, 1),
    (2, 4, 256, 384, 768, 3, 2, 2),
    (2, 2, 128, 768, 768, 3, 1, 1),
    (2, 16, 512, 192, 192, 1, 2, 1),
    (2, 8, 512, 192, 384, 1, 2, 2),
    (3, 4, 64, 64, 128, 3, 2, 2),        # ragged: Wo=32 < tile
    (1, 8, 200, 64, 64, 3, 1, 1),        # W not a multiple of the tile
]


def _conv_ref(x_nhwc, w_oihw, ks, sh, sw):
    return F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w_oihw.float(), None, (sh, sw), ks // 2)


@pytest.mark.parametrize("fwd", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("NB,H,W,Cin,Cout,ks,sh,sw", CONVS)
def test_conv_fwd_dgrad_wgrad(NB, H, W, Cin, Cout, ks, sh, sw, fwd):
    """fwd = storage format of the FORWARD tensors (x, w, y): fp16 is what the engine uses, bf16 the all-bf16 form.
    The backward GEMMs always take bf16 (tcgen05 needs one 16-bit format per MMA): bf16 copies of x and w."""
    o = ops()
    torch.manual_seed(4)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    x = torch.randn(NB, H, W, Cin, device="cuda").to(fwd)
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda") / (Cin * ks * ks) ** 0.5).to(fwd)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, ks * ks, Cin).contiguous()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, (sh, sw), ks // 2)
    rows = o.conv_stats_rows(NB, H, W, ks, sh, sw)
    stats = torch.zeros(rows, 2, Cout, device="cuda")
    y = torch.full((NB, yr.shape[2], yr.shape[3], Cout), 7.0, device="cuda", dtype=fwd)
    y = o.conv_fwd(x, wk, ks, sh, sw, y=y, stats=stats)
    assert y.shape == (NB, yr.shape[2], yr.shape[3], Cout)
    assert _rel(y.permute(0, 3, 1, 2), yr) < 1e-2
    yf = y.float()
    assert _rel(stats[:, 0].sum(0), yf.sum((0, 1, 2))) < 1e-3 or float((stats[:, 0].sum(0) - yf.sum((0, 1, 2))).abs().max()) < 0.5
    assert _rel(stats[:, 1].sum(0), (yf * yf).sum((0, 1, 2))) < 1e-3
    dy = torch.randn_like(y).bfloat16()                                       # gradients are always bf16
    if fwd != torch.bfloat16:                 # reference gradients at the bf16 copies the backward GEMMs consume
        xr = x.bfloat16().float().permute(0, 3, 1, 2).requires_grad_(True)
        wr = w.bfloat16().float().requires_grad_(True)
        yr = F.conv2d(xr, wr, None, (sh, sw), ks // 2)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    if fwd != torch.bfloat16:
        with pytest.raises(Exception):                                        # mixed bf16 x fp16 MMAs are refused
            o.conv_dgrad(dy, wk, (NB, H, W, Cin), ks, sh, sw)
        with pytest.raises(Exception):
            o.conv_wgrad_acc_t(dy, x, ks, sh, sw, torch.zeros(ks * ks, Cin, Cout, device="cuda"))
    x_fwd, wk_fwd = x, wk
    x, wk = x.bfloat16(), wk.bfloat16()                                       # the backward's bf16 copies
    dx = o.conv_dgrad(dy, wk, (NB, H, W, Cin), ks, sh, sw)                    # MN-major B operand
    assert _rel(dx.permute(0, 3, 1, 2), xr.grad) < 1e-2
    # K-major B operand from the transposed pack [Cin, taps, Cout] (CTA-pair kernel where the tile grid allows)
    packed = o.pack_weights([(w.float().contiguous(), "conv"), (w.float().contiguous(), "convT")], conv_dtype=fwd)
    assert torch.equal(packed[0], wk_fwd) and packed[0].dtype == fwd
    wt = w.float().permute(1, 2, 3, 0).reshape(Cin, ks * ks, Cout).contiguous().bfloat16()   # backward copy: bf16
    if fwd != torch.bfloat16:
        wt = wk.permute(2, 1, 0).contiguous()                                   # (bf16 of the fp16-rounded weights)
        packed[1] = wt
    assert torch.equal(packed[1], wt)
    dx2 = o.conv_dgrad(dy, wk, (NB, H, W, Cin), ks, sh, sw, w_t=wt)
    assert _rel(dx2.permute(0, 3, 1, 2), xr.grad) < 1e-2
    # transposed weight-gradient GEMM ([taps, Cin, Cout] output, (tap, channel-atom) rows, CTA pairs / SWIZZLE_64B
    # dY halves) + the one-launch unpack into OIHW (accumulating)
    gt = torch.zeros(ks * ks, Cin, Cout, device="cuda")
    o.conv_wgrad_acc_t(dy, x, ks, sh, sw, gt)
    assert _rel(gt.permute(2, 1, 0).reshape(Cout, Cin, ks, ks), wr.grad) < 1e-3
    gw2 = torch.ones(Cout, Cin, ks, ks, device="cuda")
    o.unpack_conv_grads([(gt, gw2)], transposed=True)
    assert _rel(gw2 - 1.0, wr.grad) < 1e-3
    if ks == 3 and sw == 1:
        # two-accumulator window-sharing weight gradient ([3, Cin/64, 3, 64, Cout] rows) + its unpack mode
        ga = torch.zeros(3, Cin // 64, 3, 64, Cout, device="cuda")
        o.conv_wgrad_acc_w(dy, x, sh, ga)
        want = wr.grad.permute(2, 1, 3, 0).reshape(3, Cin // 64, 64, 3, Cout).permute(0, 1, 3, 2, 4)   # kh, c, kw, 64, co
        assert _rel(ga, want) < 1e-3
        gw3 = torch.ones(Cout, Cin, ks, ks, device="cuda")
        o.unpack_conv_grads([(ga, gw3)], layout="atoms")
        assert _rel(gw3 - 1.0, wr.grad) < 1e-3
    gw = torch.zeros(Cout, Cin, ks, ks, device="cuda")
    o.conv_wgrad(dy, x, ks, sh, sw, gw, accumulate=False, transpose=True)      # K-major dY^T operand (transpose_px)
    assert _rel(gw, wr.grad) < 1e-3
    dyt = o.transpose_px(dy)
    assert torch.equal(dyt, dy.permute(0, 1, 3, 2).contiguous())
    # tap-major in-place accumulation (split-K slices reduce-add through TMA) + the multi-tensor unpack
    gt = torch.zeros(Cout, ks * ks, Cin, device="cuda")
    o.conv_wgrad_acc(dy, x, ks, sh, sw, gt)
    o.conv_wgrad_acc(dy, x, ks, sh, sw, gt)                                    # += semantics
    gw3 = torch.ones(Cout, Cin, ks, ks, device="cuda")
    o.unpack_conv_grads([(gt, gw3)])
    assert _rel(gw3 - 1, 2 * wr.grad) < 1e-3
    gw2 = torch.zeros(Cout, Cin, ks, ks, device="cuda")
    o.conv_wgrad(dy, x, ks, sh, sw, gw2, accumulate=False, transpose=False)    # both operands pixel-major
    assert _rel(gw2, wr.grad) < 1e-3


@pytest.mark.parametrize("NB,H,W,C", [(2, 8, 256, 192), (2, 4, 128, 384), (3, 2, 200, 768)])
def test_conv_dgrad_with_bn_backward_epilogue(NB, H, W, C):
    """conv2's input gradient with the BatchNorm-backward reduction of the layer in front in its epilogue (ReLU-masked
    gradient + [2][C] column sums) followed by the apply pass, against the two-pass route (conv_dgrad -> bn_bwd) and an
    fp32 torch graph (conv -> BN(train) -> ReLU -> conv)."""
    o = ops()
    torch.manual_seed(NB + C)
    Cout = C
    x = torch.randn(NB, H, W, C, device="cuda").half()                       # raw conv1 output (fp16)
    w2 = (torch.randn(Cout, C, 3, 3, device="cuda") / (3 * C ** 0.5))
    gamma = (torch.rand(C, device="cuda") + 0.5)
    beta = torch.randn(C, device="cuda") * 0.1
    P = NB * H * W
    f = x.float().view(P, C)
    part = torch.stack([f.sum(0), (f * f).sum(0)]).view(1, 2, C).contiguous()
    st = o.bn_finalize(part, P, gamma, beta, None, None, None, True)
    a1, km, a1_bf = o.bn_act_fwd(x, st, True, want_mask=True, want_bf16=True)
    dy = (torch.randn(NB, H, W, Cout, device="cuda") * 0.1).bfloat16()
    packed = o.pack_weights([(w2.contiguous(), "conv"), (w2.contiguous(), "convT")])
    wk, wt = packed[0], packed[1]
    # two-pass route
    g = o.conv_dgrad(dy, wk, (NB, H, W, C), 3, 1, 1, w_t=wt)
    dg_a, db_a = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    d_a, _, _ = o.bn_bwd(g, km, x, st, gamma, dg_a, db_a)
    # fused route
    sums = torch.zeros(3 * C, device="cuda")
    gm = o.conv_dgrad_bn(dy, wt, (NB, H, W, C), x, km, st, sums)
    assert gm is not None, "the window-sharing CTA-pair kernel serves these shapes"
    dg_b, db_b = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    d_b = o.bn_bwd_apply(gm, x, st, gamma, dg_b, db_b, sums)
    bits = ((km.view(P, C // 8, 1).int() >> torch.arange(8, device="cuda").view(1, 1, 8)) & 1).view(NB, H, W, C)
    assert torch.equal(gm.float(), g.float() * bits)                        # same GEMM, masked in the epilogue
    assert _rel(d_b, d_a) < 1e-2 and _rel(dg_b, dg_a) < 2e-3 and _rel(db_b, db_a) < 2e-3
    # fp32 torch graph
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    gam, bet = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.conv2d(F.relu(F.batch_norm(xf, None, None, gam, bet, True, 0.1, 1e-5)), w2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(d_b, xf.grad.permute(0, 2, 3, 1)) < 3e-2
    assert _rel(dg_b, gam.grad) < 2e-2 and _rel(db_b, bet.grad) < 2e-2


def test_programmatic_dependent_launch_does_not_change_results():
    """htrvt_set_pdl: the tap GEMMs' prologue may run under the tail of the kernel in front; a chain of dependent GEMMs
    (each reads what the previous one wrote) must give bit-identical results with and without it."""
    o = ops()
    torch.manual_seed(3)
    M, K = 4096, 768
    x = torch.randn(M, K, device="cuda").bfloat16()
    ws = [(torch.randn(K, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(6)]
    outs = {}
    prev = o.set_pdl(True)
    try:
        for flag in (False, True, False):
            assert o.set_pdl(flag) in (True, False)
            h = x
            for w in ws:                                   # back-to-back launches on one stream, no host sync
                y = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
                o.gemm_tn(h, w, y)
                h = y
            torch.cuda.synchronize()
            outs.setdefault(flag, []).append(h.clone())
        assert o.set_pdl(True) is False                    # returns the previous setting
    finally:
        o.set_pdl(prev)
    assert torch.equal(outs[False][0], outs[False][1])
    assert torch.equal(outs[False][0], outs[True][0])
    assert torch.isfinite(outs[True][0].float()).all()
