"""Thin torch-tensor -> C-ABI call layer (raw device pointers, current CUDA stream).

PyTorch is used here for device memory and streams only; all arithmetic runs in csrc/*.cu.
"""
import ctypes

import torch

from ._lib import HtrvtError, check, lib

EPI_BF16, EPI_BIAS, EPI_GELU, EPI_RESID, EPI_ACCUM, EPI_STATS, EPI_QKV, EPI_RELU = (1 << i for i in range(8))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise HtrvtError("htr-vt_b200 kernels need CUDA tensors (there is no CPU fallback)")


_ws_cache = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only per-device scratch buffer (owned by PyTorch's caching allocator)."""
    key = (device.type, device.index)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------
# GEMMs
# ------------------------------------------------------------------------------------------------
def gemm_tn(x, w, out, *, bias=None, resid=None, out2=None, flags=0, alpha=1.0, qkv=None):
    """out[M,N] = epilogue(alpha * x[M,K] @ w[N,K]^T).  x, w bf16 row-major."""
    _need_cuda(x, w, out)
    M, K = x.shape
    N = w.shape[0]
    f = flags | (EPI_BF16 if out.dtype == torch.bfloat16 else 0) | (EPI_BIAS if bias is not None else 0) \
        | (EPI_RESID if resid is not None else 0) | (EPI_GELU if out2 is not None else 0)
    qb, qt, qh, qd = qkv if qkv is not None else (0, 0, 0, 0)
    if qkv is not None:
        f |= EPI_QKV
    ldo = out.stride(0) if (qkv is None and out.dim() == 2) else 0
    check(lib().htrvt_gemm_tn(_p(x), x.stride(0), _p(w), w.stride(0), M, N, K, f, _p(bias), _p(resid), _p(out), ldo,
                              _p(out2), alpha, qb, qt, qh, qd, _stream()), "htrvt_gemm_tn")
    return out


def gemm_nn(dy, w, out, *, accumulate=False, alpha=1.0):
    """out[M,N] (+)= dy[M,K] @ w[K,N]   (w row-major [K,N]: the nn.Linear weight itself for dgrad)."""
    _need_cuda(dy, w, out)
    M, K = dy.shape
    N = w.shape[1]
    f = (EPI_BF16 if out.dtype == torch.bfloat16 else 0) | (EPI_ACCUM if accumulate else 0)
    check(lib().htrvt_gemm_nn(_p(dy), dy.stride(0), _p(w), w.stride(0), M, N, K, f, _p(out), out.stride(0), alpha,
                              _stream()), "htrvt_gemm_nn")
    return out


def linear_wgrad(dy, x, grad, *, accumulate=True):
    """grad[N,K] (+)= dy[M,N]^T @ x[M,K]  (fp32 grad)."""
    _need_cuda(dy, x, grad)
    M, N = dy.shape
    K = x.shape[1]
    nbytes = lib().htrvt_wgrad_workspace_bytes(N, K, 1, M)
    ws = workspace(nbytes, dy.device)
    check(lib().htrvt_linear_wgrad(_p(dy), dy.stride(0), _p(x), x.stride(0), M, N, K, _p(grad), int(accumulate),
                                   _p(ws), ws.numel(), _stream()), "htrvt_linear_wgrad")
    return grad


def conv_out_hw(H, W, ks, sh, sw):
    pad = ks // 2
    return (H + 2 * pad - ks) // sh + 1, (W + 2 * pad - ks) // sw + 1


def conv_fwd(x, w, ks, sh, sw, y=None, stats=None, relu=False):
    """x [N,H,W,Cin] bf16 NHWC, w [Cout, ks*ks, Cin] bf16 -> y [N,Ho,Wo,Cout] bf16 (raw conv output).
    stats: optional fp32 [rows, 2, Cout] per-tile column sum / sum-of-squares partials."""
    _need_cuda(x, w)
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    Ho, Wo = conv_out_hw(H, W, ks, sh, sw)
    if y is None:
        y = torch.empty((N, Ho, Wo, Cout), dtype=torch.bfloat16, device=x.device)
    check(lib().htrvt_conv_fwd(_p(x), N, H, W, Cin, _p(w), Cout, ks, sh, sw, _p(y), _p(stats),
                               EPI_RELU if relu else 0, _stream()), "htrvt_conv_fwd")
    return y


def conv_stats_rows(N, H, W, ks, sh, sw):
    return lib().htrvt_conv_fwd_stats_rows(N, H, W, ks, sh, sw)


def conv_dgrad(dy, w, x_shape, ks, sh, sw, dx=None, accumulate=False):
    _need_cuda(dy, w)
    N, H, W, Cin = x_shape
    Cout = w.shape[0]
    if dx is None:
        dx = torch.zeros(x_shape, dtype=torch.bfloat16, device=dy.device) if (ks == 1 and (sh > 1 or sw > 1)) \
            else torch.empty(x_shape, dtype=torch.bfloat16, device=dy.device)
    check(lib().htrvt_conv_dgrad(_p(dy), N, H, W, Cin, _p(w), Cout, ks, sh, sw, _p(dx), int(accumulate), _stream()),
          "htrvt_conv_dgrad")
    return dx


def conv_wgrad(dy, x, ks, sh, sw, grad_oihw, accumulate=True):
    _need_cuda(dy, x, grad_oihw)
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    Ho, Wo = conv_out_hw(H, W, ks, sh, sw)
    nbytes = lib().htrvt_wgrad_workspace_bytes(Cout, Cin, ks * ks, Wo)  # chunks are per row; bound below
    nbytes = max(nbytes, 64 * Cout * ks * ks * Cin * 4)
    ws = workspace(nbytes, dy.device)
    check(lib().htrvt_conv_wgrad(_p(dy), _p(x), N, H, W, Cin, Cout, ks, sh, sw, _p(grad_oihw), int(accumulate),
                                 _p(ws), ws.numel(), _stream()), "htrvt_conv_wgrad")
    return grad_oihw


# ------------------------------------------------------------------------------------------------
# CTC + decode
# ------------------------------------------------------------------------------------------------
def ctc_loss_grad(x, targets, input_lengths, target_lengths, *, layout, is_logprob, want_grad=True,
                  max_target_len=-1, grad_scale=None, grad_scale_const=1.0):
    """x: fp32 logits [B,T,C] (layout 'btc') or log-probs [T,B,C] (layout 'tbc').
    Returns (nll [B] fp32, grad with the shape of x or None)."""
    _need_cuda(x, targets, target_lengths)
    if x.dtype != torch.float32 or x.stride(-1) != 1:
        raise HtrvtError("ctc_loss_grad expects fp32 input with a contiguous class axis")
    if layout == "btc":
        B, T, C = x.shape
        sb, st = x.stride(0), x.stride(1)
    else:
        T, B, C = x.shape
        sb, st = x.stride(1), x.stride(0)
    dev = x.device
    tg = targets.to(device=dev, dtype=torch.int32)
    tl = target_lengths.to(device=dev, dtype=torch.int32).contiguous()
    il = None if input_lengths is None else input_lengths.to(device=dev, dtype=torch.int32).contiguous()
    if tg.dim() == 2:
        tgt_stride = tg.stride(0)
        tg = tg.contiguous()
        tgt_stride = tg.stride(0)
    else:
        tg = tg.contiguous()
        tgt_stride = 0
    nll = torch.empty(B, dtype=torch.float32, device=dev)
    grad = torch.empty_like(x) if want_grad else None
    if grad is not None:
        gsb, gst = (grad.stride(0), grad.stride(1)) if layout == "btc" else (grad.stride(1), grad.stride(0))
    else:
        gsb = gst = 0
    nbytes = lib().htrvt_ctc_workspace_bytes(B, T, C, max_target_len)
    ws = workspace(nbytes, dev) if nbytes else None
    check(lib().htrvt_ctc_loss_grad(_p(x), sb, st, int(is_logprob), _p(tg), tgt_stride, _p(il), _p(tl), B, T, C,
                                    max_target_len, _p(nll), _p(grad), gsb, gst, _p(grad_scale), grad_scale_const,
                                    _p(ws), ws.numel() if ws is not None else 0, _stream()), "htrvt_ctc_loss_grad")
    return nll, grad


def greedy_decode_ids(logits, n_character, lengths=None, layout="btc", want_raw=False):
    """argmax + collapse on device.  Returns (ids [B,T] int32, lens [B] int32, raw [B,T] int32 | None)."""
    _need_cuda(logits)
    if logits.dtype != torch.float32 or logits.stride(-1) != 1:
        raise HtrvtError("greedy_decode expects fp32 logits with a contiguous class axis")
    if layout == "btc":
        B, T, C = logits.shape
        sb, st = logits.stride(0), logits.stride(1)
    else:
        T, B, C = logits.shape
        sb, st = logits.stride(1), logits.stride(0)
    dev = logits.device
    ids = torch.empty((B, T), dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    raw = torch.empty((B, T), dtype=torch.int32, device=dev) if want_raw else None
    ln = None if lengths is None else lengths.to(device=dev, dtype=torch.int32).contiguous()
    check(lib().htrvt_greedy_decode(_p(logits), sb, st, B, T, C, _p(ln), n_character, _p(ids), _p(lens), _p(raw),
                                    _stream()), "htrvt_greedy_decode")
    return ids, lens, raw


def ctc_collapse(index_flat, lengths, n_character):
    """Collapse a sample-major index stream (reference decode() input).  -> (ids [B,Tmax], lens [B])."""
    _need_cuda(index_flat)
    dev = index_flat.device
    ln_host = lengths.detach().to("cpu", torch.int64)
    B = int(ln_host.numel())
    Tmax = int(ln_host.max()) if B else 0
    ln = ln_host.to(device=dev, dtype=torch.int32)
    idx = index_flat.contiguous()
    if idx.dtype not in (torch.int64, torch.int32):
        idx = idx.to(torch.int64)
    ids = torch.empty((B, max(Tmax, 1)), dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    check(lib().htrvt_ctc_collapse(_p(idx), int(idx.dtype == torch.int64), _p(ln), B, max(Tmax, 1), n_character,
                                   _p(ids), _p(lens), _stream()), "htrvt_ctc_collapse")
    return ids, lens
