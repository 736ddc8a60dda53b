"""Exponential moving average of the model weights on one multi-tensor kernel launch.

The reference keeps its EMA in `ModelEma` (model_v1/utils/utils.py:127-173): a frozen deep copy of the model whose
state_dict entries follow `ema = d * ema + (1 - d) * model` after every optimizer step, one mul/add/copy_ triple per
entry (150 entries in model_v1).  Two ways to get the fused update:

  * `install_fused_update(cls)` swaps ONLY the `update` method of the reference's own `ModelEma` class (nothing of the
    reference is restated; checkpoints, `.ema`, `.decay` keep working):
        from utils import utils; install_fused_update(utils.ModelEma)
  * `ModelEma` below: a stand-alone class with the same constructor arguments, attributes and methods for callers that
    do not import the reference's utils module at all.

Semantics kept: every state_dict entry is averaged, BatchNorm running statistics included; the int64
`num_batches_tracked` counters go through the reference's own expression (float math, truncating copy back); the warm-up
decay is min(decay, (1 + n) / (10 + n)) when `num_updates` is given.
"""
import copy

import torch

from .. import ops as _ops
from .._lib import check, lib
from .sam import _numels, _ptrs, _stream


def effective_decay(decay, num_updates=-1):
    """model_v1/utils/utils.py:160-163."""
    if num_updates < 0:
        return decay
    return min(decay, (1 + num_updates) / (10 + num_updates))


def _fusable(e, m):
    return (e.is_cuda and m.is_cuda and e.device == m.device and e.dtype == torch.float32 and m.dtype == torch.float32
            and e.is_contiguous() and m.is_contiguous())


@torch.no_grad()
def ema_update_(ema_module, model, decay, device="", key_prefix=""):
    """ema_module.state_dict()[k] <- decay * itself + (1 - decay) * model.state_dict()[key_prefix + k], in place.
    fp32 CUDA entries go through htrvt_mt_ema (<= 192 tensors per launch); anything else uses the reference arithmetic."""
    src = model.state_dict()
    pairs_e, pairs_m, rest = [], [], {}
    for name, e in ema_module.state_dict().items():
        m = src[key_prefix + name].detach()
        if device:
            m = m.to(device=device)
        if _fusable(e, m):
            pairs_e.append(e)
            pairs_m.append(m)
        else:
            rest.setdefault((e.dtype, m.dtype, tuple(e.shape), tuple(m.shape), e.device, m.device), []).append((e, m))
    for (_, _, es, ms_, ed, md), group in rest.items():
        if len(group) > 1 and es == ms_ and ed == md and ed.type == "cuda":
            # the int64 `num_batches_tracked` counters (one per BatchNorm): the reference's expression on the stacked
            # entries - same type promotion (an integer tensor times a Python float is float32 whether 0-dim or not),
            # same truncating conversion on the copy back - in 6 launches instead of 4 per counter
            E = torch.stack([e for e, _ in group])
            M = torch.stack([m for _, m in group])
            R = (E * decay + (1. - decay) * M).to(E.dtype)
            if hasattr(torch, "_foreach_copy_"):
                torch._foreach_copy_([e for e, _ in group], list(R.unbind(0)))
            else:
                for (e, _), r in zip(group, R.unbind(0)):
                    e.copy_(r)
        else:
            for e, m in group:
                e.copy_(e * decay + (1. - decay) * m)
    if pairs_e:
        check(lib().htrvt_mt_ema(len(pairs_e), _ptrs(pairs_e), _ptrs(pairs_m), _numels(pairs_e), float(decay),
                                 _stream()), "htrvt_mt_ema")
        _ops.weights_changed()           # raw-pointer writes: the engine's packed-weight cache keys on this epoch


def _fused_update(self, model, num_updates=-1):
    prefix = "module." if (hasattr(model, "module") and not self.ema_has_module) else ""
    ema_update_(self.ema, model, effective_decay(self.decay, num_updates), self.device, prefix)


def install_fused_update(ema_cls):
    """Replace `ema_cls.update` (the reference's ModelEma, or any class with .ema/.decay/.device/.ema_has_module)."""
    ema_cls.update = _fused_update
    return ema_cls


class ModelEma(object):
    """Stand-alone EMA holder, interface of model_v1/utils/utils.py:127-173 (ctor args, .ema, .decay, .device,
    .ema_has_module, update(model, num_updates), resume from a checkpoint's 'state_dict_ema')."""

    def __init__(self, model, decay=0.9999, device='', resume=''):
        shadow = copy.deepcopy(model).eval()
        if device:
            shadow.to(device=device)
        self.ema, self.decay, self.device = shadow, decay, device
        self.ema_has_module = hasattr(shadow, 'module')
        if resume:
            self._load_checkpoint(resume)
        shadow.requires_grad_(False)

    def _load_checkpoint(self, checkpoint_path, mapl=None):
        from .checkpoint import strip_module_prefix
        ckpt = torch.load(checkpoint_path, map_location=mapl, weights_only=False)
        weights = ckpt.get('state_dict_ema') if isinstance(ckpt, dict) else None
        if weights is None:
            print("=> Failed to find state_dict_ema, starting from loaded model weights")
            return
        weights = strip_module_prefix(weights)
        if self.ema_has_module:
            weights = {'module.' + k: v for k, v in weights.items()}
        self.ema.load_state_dict(weights)
        print("=> Loaded state_dict_ema")

    update = _fused_update
