"""`SAM(params, base_optimizer, rho, adaptive, **kwargs)` with the reference's interface (model_v1/utils/sam.py) on
multi-tensor sm_100a kernels (csrc/optim.cu).

The reference walks the parameter list in Python: a `norm` launch per tensor + `stack` + `norm` for the gradient
norm, a `clone` and an `add_` per tensor in `first_step`, a pointer swap per tensor and the base optimizer's own
per-tensor (or foreach) update in `second_step` - ~10^3 launches per iteration for the 102 parameters of HTR-VT.
Here each pass (gradient norm, climb, restore + AdamW) is one launch per <= 192 tensors and the norm never visits the
host.  Same call sequence as model_v1/train.py:117-126: `first_step(zero_grad=True)`, forward/backward,
`second_step(zero_grad=True)`; `param_groups` (lr set by utils.update_lr_cos), `state_dict()` and `zero_grad()` behave
as in the reference.  Only plain AdamW (the optimizer the reference trains with, train.py:93) has a fused path; any other
base optimizer - or AdamW with amsgrad / maximize in any group - takes the reference sequence: restore, then the base
optimizer's own `.step()`.
"""
import ctypes

import torch

from .. import ops as _ops
from .._lib import check, lib


def _ptrs(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _numels(tensors):
    arr = (ctypes.c_longlong * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.numel()
    return arr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dense_f32(t):
    return t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()


class SAM(torch.optim.Optimizer):
    def __init__(self, params, base_optimizer, rho=0.05, adaptive=False, **kwargs):
        assert rho >= 0.0, f"Invalid rho, should be non-negative: {rho}"
        defaults = dict(rho=rho, adaptive=adaptive, **kwargs)
        super(SAM, self).__init__(params, defaults)
        self.base_optimizer = base_optimizer(self.param_groups, **kwargs)
        self.param_groups = self.base_optimizer.param_groups
        self.defaults.update(self.base_optimizer.defaults)
        self._fused_adamw = isinstance(self.base_optimizer, torch.optim.AdamW)
        self._norm2 = None

    # -- helpers ------------------------------------------------------------------------------------
    def _live(self, group):
        """(parameters with a gradient, their gradients) - each `.grad` read once per pass: this runs on the host for
        every one of the ~100 tensors, four times per iteration."""
        ps, gs = [], []
        for p in group["params"]:
            g = p.grad
            if g is None:
                continue
            if not (_dense_f32(p) and _dense_f32(g)):
                raise RuntimeError("htr-vt_b200 SAM needs contiguous fp32 CUDA parameters and gradients")
            ps.append(p)
            gs.append(g)
        return ps, gs

    def _grad_norm_device(self, live=None):
        """sum |g|^2 (adaptive: |abs(w) g|^2) over every parameter, left on the device as a float64 scalar."""
        if live is None:
            live = [(group,) + self._live(group) for group in self.param_groups]
        dev = self.param_groups[0]["params"][0].device
        if self._norm2 is None or self._norm2.device != dev:
            self._norm2 = torch.zeros((), dtype=torch.float64, device=dev)
        self._norm2.zero_()
        for group, ps, gs in live:
            if not ps:
                continue
            check(lib().htrvt_mt_sqnorm(len(ps), _ptrs(gs), _ptrs(ps), _numels(ps),
                                        int(bool(group["adaptive"])), ctypes.c_void_p(self._norm2.data_ptr()),
                                        _stream()), "htrvt_mt_sqnorm")
        return self._norm2

    def _grad_norm(self):                     # reference name (sam.py:49): returns the norm as a tensor
        return self._grad_norm_device().sqrt().float()

    # -- the reference's two steps --------------------------------------------------------------------
    @torch.no_grad()
    def first_step(self, zero_grad=False):
        live = [(group,) + self._live(group) for group in self.param_groups]
        norm2 = self._grad_norm_device(live)
        for group, ps, gs in live:
            if not ps:
                continue
            olds = []
            for p in ps:
                st = self.state[p]
                if "old_p" not in st or st["old_p"].shape != p.shape:
                    st["old_p"] = torch.empty_like(p.data)
                olds.append(st["old_p"])
            check(lib().htrvt_mt_sam_first(len(ps), _ptrs(ps), _ptrs(gs), _ptrs(olds), _numels(ps), ctypes.c_void_p(norm2.data_ptr()),
                                           float(group["rho"]), int(bool(group["adaptive"])), _stream()),
                  "htrvt_mt_sam_first")
        _ops.weights_changed()
        if zero_grad:
            self.zero_grad()

    def _group_fusable(self, group):
        """The fused restore + AdamW kernel implements plain decoupled-weight-decay Adam only."""
        return (self._fused_adamw and not group.get("amsgrad", False) and not group.get("maximize", False)
                and not group.get("capturable", False) and not group.get("differentiable", False))

    @torch.no_grad()
    def second_step(self, zero_grad=False):
        live = [(group,) + self._live(group) for group in self.param_groups]
        if not all(self._group_fusable(g) for g, ps, _ in live if ps):
            # any group the fused kernel cannot express (other base optimizer, amsgrad / maximize, ...): the reference
            # sequence for EVERY group - back to "w" from "w + e(w)", then the base optimizer's own update (sam.py:31-37)
            for _, ps, _ in live:
                for p in ps:
                    p.data.copy_(self.state[p]["old_p"])
            self.base_optimizer.step()
            _ops.weights_changed()
            if zero_grad:
                self.zero_grad()
            return
        bst = self.base_optimizer.state
        for group, ps, gs in live:
            if not ps:
                continue
            olds = [self.state[p]["old_p"] for p in ps]
            ms, vs, counters = [], [], []
            for p in ps:
                st = bst[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p.data)
                    st["exp_avg_sq"] = torch.zeros_like(p.data)
                counters.append(st["step"])
                ms.append(st["exp_avg"])
                vs.append(st["exp_avg_sq"])
            torch._foreach_add_(counters, 1)          # torch.optim.AdamW's per-parameter step tensors, one call
            steps = {int(c) for c in counters}
            if len(steps) != 1:
                raise RuntimeError("htr-vt_b200 SAM: parameters of one group must share their step count")
            b1, b2 = group["betas"]
            check(lib().htrvt_mt_adamw(len(ps), _ptrs(ps), _ptrs(gs), _ptrs(ms),
                                       _ptrs(vs), _ptrs(olds), _numels(ps), float(group["lr"]), float(b1), float(b2),
                                       float(group["eps"]), float(group["weight_decay"]), steps.pop(), _stream()),
                  "htrvt_mt_adamw")
        _ops.weights_changed()
        if zero_grad:
            self.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is not None, "Sharpness Aware Minimization requires closure, but it was not provided"
        closure = torch.enable_grad()(closure)
        self.first_step(zero_grad=True)
        closure()
        self.second_step()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.base_optimizer.param_groups = self.param_groups
