"""`ModelEma` with the reference's interface (model_v1/utils/utils.py:127-173) on one multi-tensor launch per <= 48
state_dict entries instead of a `mul`/`add`/`copy_` triple per entry (150 entries in model_v1).  Integer entries
(`num_batches_tracked`) keep the reference arithmetic, `ema_v.copy_(ema_v * d + (1 - d) * model_v)`, on the torch path."""
import ctypes
from collections import OrderedDict
from copy import deepcopy

import torch

from .. import ops as _ops
from .._lib import check, lib
from .sam import _numels, _ptrs, _stream


class ModelEma:
    def __init__(self, model, decay=0.9999, device='', resume=''):
        self.ema = deepcopy(model)
        self.ema.eval()
        self.decay = decay
        self.device = device
        if device:
            self.ema.to(device=device)
        self.ema_has_module = hasattr(self.ema, 'module')
        if resume:
            self._load_checkpoint(resume)
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def _load_checkpoint(self, checkpoint_path, mapl=None):
        checkpoint = torch.load(checkpoint_path, map_location=mapl)
        assert isinstance(checkpoint, dict)
        if 'state_dict_ema' in checkpoint:
            new_state_dict = OrderedDict()
            for k, v in checkpoint['state_dict_ema'].items():
                if self.ema_has_module:
                    name = 'module.' + k if not k.startswith('module') else k
                else:
                    name = k
                new_state_dict[name] = v
            self.ema.load_state_dict(new_state_dict)
            print("=> Loaded state_dict_ema")
        else:
            print("=> Failed to find state_dict_ema, starting from loaded model weights")

    def update(self, model, num_updates=-1):
        needs_module = hasattr(model, 'module') and not self.ema_has_module
        if num_updates >= 0:
            _cdecay = min(self.decay, (1 + num_updates) / (10 + num_updates))
        else:
            _cdecay = self.decay
        with torch.no_grad():
            msd = model.state_dict()
            fused_e, fused_m = [], []
            for k, ema_v in self.ema.state_dict().items():
                if needs_module:
                    k = 'module.' + k
                model_v = msd[k].detach()
                if self.device:
                    model_v = model_v.to(device=self.device)
                if (ema_v.is_cuda and model_v.is_cuda and ema_v.dtype == torch.float32 and model_v.dtype == torch.float32
                        and ema_v.is_contiguous() and model_v.is_contiguous() and ema_v.device == model_v.device):
                    fused_e.append(ema_v)
                    fused_m.append(model_v)
                else:
                    ema_v.copy_(ema_v * _cdecay + (1. - _cdecay) * model_v)
            if fused_e:
                check(lib().htrvt_mt_ema(len(fused_e), _ptrs(fused_e), _ptrs(fused_m), _numels(fused_e),
                                         float(_cdecay), _stream()), "htrvt_mt_ema")
                _ops.weights_changed()
