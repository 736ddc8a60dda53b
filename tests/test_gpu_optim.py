"""GPU parity of the multi-tensor SAM(AdamW) / EMA passes against the reference algorithm (model_v1/utils/sam.py,
torch.optim.AdamW, model_v1/utils/utils.py:158-173) restated with plain torch ops on the same tensors."""
import copy
from importlib import import_module

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    import htrvt_b200  # noqa: F401
    return import_module("htr-vt_b200.utils.sam"), import_module("htr-vt_b200.utils.utils")


class _RefSAM(torch.optim.Optimizer):
    """The reference SAM, verbatim semantics (model_v1/utils/sam.py:15-59)."""

    def __init__(self, params, base_optimizer, rho=0.05, adaptive=False, **kwargs):
        super().__init__(params, dict(rho=rho, adaptive=adaptive, **kwargs))
        self.base_optimizer = base_optimizer(self.param_groups, **kwargs)
        self.param_groups = self.base_optimizer.param_groups

    @torch.no_grad()
    def first_step(self):
        norm = torch.norm(torch.stack([((torch.abs(p) if g["adaptive"] else 1.0) * p.grad).norm(p=2)
                                       for g in self.param_groups for p in g["params"] if p.grad is not None]), p=2)
        for group in self.param_groups:
            scale = group["rho"] / (norm + 1e-12)
            for p in group["params"]:
                if p.grad is None:
                    continue
                self.state[p]["old_p"] = p.data.clone()
                p.add_((torch.pow(p, 2) if group["adaptive"] else 1.0) * p.grad * scale.to(p))

    @torch.no_grad()
    def second_step(self):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.data = self.state[p]["old_p"]
        self.base_optimizer.step()


@pytest.mark.parametrize("adaptive", [False, True])
def test_sam_adamw_matches_reference(adaptive):
    sam, _ = _mods()
    torch.manual_seed(0)
    shapes = [(768, 3072), (3072,), (192, 1, 3, 3), (1, 1, 768), (70001,), (5,)] + [(64, 64)] * 60   # > 48 tensors
    base = [torch.randn(s, device="cuda") for s in shapes]
    ours = [torch.nn.Parameter(t.clone()) for t in base] + [torch.nn.Parameter(torch.randn(7, device="cuda"))]  # last: no grad
    ref = [torch.nn.Parameter(t.clone()) for t in base] + [torch.nn.Parameter(ours[-1].detach().clone())]
    kw = dict(lr=3e-3, betas=(0.9, 0.99), weight_decay=0.5)
    o = sam.SAM(ours, torch.optim.AdamW, rho=0.05, adaptive=adaptive, **kw)
    r = _RefSAM(ref, torch.optim.AdamW, rho=0.05, adaptive=adaptive, **kw)
    for it in range(3):
        for g in o.param_groups:
            g["lr"] = 3e-3 * (it + 1)               # utils.update_lr_cos rewrites param_group['lr'] every iteration
        for g in r.param_groups:
            g["lr"] = 3e-3 * (it + 1)
        g1 = [torch.randn_like(t) for t in base]
        g2 = [torch.randn_like(t) for t in base]
        for ps in (ours, ref):
            for p, g in zip(ps, g1):
                p.grad = g.clone()
        o.first_step(zero_grad=True)
        r.first_step()
        for a, b in zip(ours[:-1], ref[:-1]):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        assert ours[0].grad is None                       # zero_grad=True
        for ps in (ours, ref):
            for p, g in zip(ps, g2):
                p.grad = g.clone()
        o.second_step(zero_grad=True)
        r.second_step()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), float((a - b).abs().max())
    assert torch.equal(ours[-1], ref[-1])                 # parameter without gradient untouched
    st = o.base_optimizer.state[ours[0]]
    assert int(st["step"]) == 3 and st["exp_avg"].shape == ours[0].shape
    assert float((o._grad_norm() - 0).abs()) >= 0         # reference helper name kept


def test_model_ema_matches_reference():
    _, U = _mods()
    H = import_module("htr-vt_b200.model.HTR_VT")
    torch.manual_seed(1)
    m = H.create_model(80, [64, 512]).cuda()
    ema = U.ModelEma(m, decay=0.99)
    want = {k: v.clone() for k, v in ema.ema.state_dict().items()}
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        m.patch_embed.bn1.running_mean.add_(1.0)
        m.patch_embed.bn1.num_batches_tracked.add_(5)
    for upd in (0, 7):
        d = min(0.99, (1 + upd) / (10 + upd))
        msd = m.state_dict()
        for k in want:
            want[k] = (want[k] * d + (1.0 - d) * msd[k]).to(want[k].dtype)        # utils.py:173 incl. the int64 cast
        ema.update(m, num_updates=upd)
    got = ema.ema.state_dict()
    assert list(got.keys()) == list(m.state_dict().keys())
    for k in want:
        if want[k].dtype.is_floating_point:
            assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-7), k
        else:
            assert torch.equal(got[k], want[k]), k
    assert not any(p.requires_grad for p in ema.ema.parameters())
    assert copy.deepcopy(ema.ema) is not None


@pytest.mark.parametrize("kw", [dict(amsgrad=True), dict(maximize=True)])
def test_sam_unfusable_adamw_flags_still_update(kw):
    """AdamW options the fused kernel does not express take the reference sequence (restore + base_optimizer.step());
    the weights must move exactly as the reference's do (they used to be silently skipped)."""
    sam, _ = _mods()
    torch.manual_seed(2)
    base = [torch.randn(33, 17, device="cuda"), torch.randn(129, device="cuda")]
    ours = [torch.nn.Parameter(t.clone()) for t in base]
    ref = [torch.nn.Parameter(t.clone()) for t in base]
    o = sam.SAM(ours, torch.optim.AdamW, rho=0.05, lr=1e-2, weight_decay=0.1, **kw)
    r = _RefSAM(ref, torch.optim.AdamW, rho=0.05, lr=1e-2, weight_decay=0.1, **kw)
    for _ in range(2):
        g1 = [torch.randn_like(t) for t in base]
        g2 = [torch.randn_like(t) for t in base]
        for ps in (ours, ref):
            for p, g in zip(ps, g1):
                p.grad = g.clone()
        o.first_step(zero_grad=True)
        r.first_step()
        for ps in (ours, ref):
            for p, g in zip(ps, g2):
                p.grad = g.clone()
        o.second_step(zero_grad=True)
        r.second_step()
    for a, b, t in zip(ours, ref, base):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        assert float((a - t).abs().max()) > 1e-3           # the weights did change


def test_install_fused_update_on_a_foreign_ema_class():
    """`install_fused_update` swaps only `update` on a class with the reference's attributes (the integration route
    that restates nothing of model_v1/utils/utils.py)."""
    _, U = _mods()

    class ForeignEma(object):
        def __init__(self, model, decay):
            self.ema = copy.deepcopy(model).eval()
            self.decay, self.device, self.ema_has_module = decay, '', False

        def update(self, model, num_updates=-1):
            raise AssertionError("should have been replaced")

    U.install_fused_update(ForeignEma)
    m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8)).cuda()
    e = ForeignEma(m, 0.5)
    before = {k: v.clone() for k, v in e.ema.state_dict().items()}
    with torch.no_grad():
        m[0].weight.add_(1.0)
        m[1].num_batches_tracked.add_(4)
    e.update(m)
    got = e.ema.state_dict()
    assert torch.allclose(got["0.weight"], before["0.weight"] + 0.5)
    assert int(got["1.num_batches_tracked"]) == 2
