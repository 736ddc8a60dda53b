"""CTC loss / greedy decode with the reference's call signatures.

`CTCLoss` mirrors `torch.nn.CTCLoss(reduction='none', zero_infinity=True)` as the reference builds it
(model_v1/train.py:95, test.py:56) and calls it (train.py:27-29, valid.py:36-38):
    criterion(log_probs[T,B,C] fp32 (log_softmax'd), targets 1-D int32 concatenated,
              input_lengths int32 [B] (CPU or CUDA), target_lengths int32 [B]) -> nll [B]
The gradient handed back is ATen's convention (softmax - posterior) * grad_out.
`ctc_loss_from_logits` is the fused form (log-softmax folded into the kernel, no [T,B,C] copy).
"""
import torch

from . import ops


class _CtcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, targets, input_lengths, target_lengths, layout, is_logprob, max_target_len):
        need = x.requires_grad
        nll, grad = ops.ctc_loss_grad(x.detach(), targets, input_lengths, target_lengths, layout=layout,
                                      is_logprob=is_logprob, want_grad=need, max_target_len=max_target_len)
        ctx.layout = layout
        if need:
            ctx.save_for_backward(grad)
        return nll

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        g = gout.to(grad.dtype)
        g = g.view(-1, 1, 1) if ctx.layout == "btc" else g.view(1, -1, 1)
        return grad * g, None, None, None, None, None, None


def _host_max(target_lengths):
    return int(target_lengths.max()) if (target_lengths.device.type == "cpu" and target_lengths.numel()) else -1


class CTCLoss(torch.nn.Module):
    """Drop-in for torch.nn.CTCLoss as configured by the reference (blank=0)."""

    def __init__(self, blank: int = 0, reduction: str = "none", zero_infinity: bool = True):
        super().__init__()
        if blank != 0:
            raise ValueError("htr-vt_b200 CTCLoss implements blank=0 (the reference's convention)")
        if not zero_infinity:
            raise ValueError("htr-vt_b200 CTCLoss implements zero_infinity=True (model_v1/train.py:95)")
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(reduction)
        self.reduction = reduction

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        if log_probs.dim() != 3:
            raise ValueError("log_probs must be [T, B, C]")
        x = log_probs if log_probs.dtype == torch.float32 else log_probs.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        nll = _CtcFn.apply(x, targets, input_lengths, target_lengths, "tbc", True, _host_max(target_lengths))
        if self.reduction == "none":
            return nll
        if self.reduction == "sum":
            return nll.sum()
        tl = target_lengths.to(nll.device).clamp_min(1).to(nll.dtype)
        return (nll / tl).mean()


def ctc_loss_from_logits(logits, targets, target_lengths, input_lengths=None, max_target_len=None):
    """Fused log_softmax + CTC on logits [B,T,C] (fp32).  Returns per-sample nll [B]."""
    x = logits if logits.dtype == torch.float32 else logits.float()
    if x.stride(-1) != 1:
        x = x.contiguous()
    mtl = _host_max(target_lengths) if max_target_len is None else int(max_target_len)
    return _CtcFn.apply(x, targets, input_lengths, target_lengths, "btc", False, mtl)


def greedy_decode(logits, n_character, lengths=None):
    """logits [B,T,C] fp32 -> (ids [B,T] int32 compacted, lens [B] int32), all on device."""
    ids, lens, _ = ops.greedy_decode_ids(logits if logits.dtype == torch.float32 else logits.float(),
                                         n_character, lengths)
    return ids, lens
