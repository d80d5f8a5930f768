"""GPU: FusedAdam (d2gs_adam_step) against torch.optim.Adam — the optimiser the reference itself uses
(scene/gaussian_model.py:203, lr=0.0 overridden per group, eps=1e-15) — and the fused densification statistics."""
import time

import numpy as np
import pytest
import torch

import util

pytestmark = pytest.mark.gpu


def _groups(dev, seed, shapes):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]
    lrs = [0.00016 * 5, 0.0025, 0.0025 / 20.0, 0.05, 0.005 * 5, 0.001, 0.0025]
    return ps, [{"params": [p], "lr": lrs[i % len(lrs)], "name": f"g{i}"} for i, p in enumerate(ps)]


def test_fused_adam_tracks_torch_adam(cuda_device):
    from d2gs_b200.optim import FusedAdam
    dev = cuda_device
    shapes = [(5000, 3), (5000, 1, 3), (5000, 15, 3), (5000, 1), (5000, 2), (5000, 4), (5000, 8), (256, 93), (256,), (13, 256), (1,), (7,)]
    pa, ga = _groups(dev, 0, shapes)
    pb, gb = _groups(dev, 0, shapes)
    ref = torch.optim.Adam(ga, lr=0.0, eps=1e-15)
    ours = FusedAdam(gb, lr=0.0, eps=1e-15)
    gen = torch.Generator().manual_seed(1)
    for step in range(25):
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, generator=gen) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=gen)))
            if step % 5 == 3 and a.numel() > 100:
                g[: a.shape[0] // 2] = 0.0                      # invisible surfels: zero gradient, the moments still decay
            a.grad = g.to(dev); b.grad = g.to(dev).clone()
        if step == 10:                                          # the xyz learning-rate schedule changes a group's lr
            ga[0]["lr"] = gb[0]["lr"] = 3e-4
        if step == 12:                                          # a parameter without gradient is skipped by both
            pa[3].grad = None; pb[3].grad = None
        ref.step(); ours.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert util.rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 1e-6
        sa, sb = ref.state[a], ours.state[b]
        assert float(sa["step"]) == float(sb["step"])
        assert util.rel_err(sb["exp_avg"].cpu().numpy(), sa["exp_avg"].cpu().numpy()) < 1e-6
        assert util.rel_err(sb["exp_avg_sq"].cpu().numpy(), sa["exp_avg_sq"].cpu().numpy()) < 1e-6
    # state_dict round trip through the torch class (same layout)
    sd = ours.state_dict()
    again = torch.optim.Adam(gb, lr=0.0, eps=1e-15)
    again.load_state_dict(sd)
    assert float(again.state[pb[0]]["step"]) == 25.0


def test_fused_adam_survives_the_reference_densification_edits(cuda_device):
    """cat_tensors_to_optimizer / _prune_optimizer (scene/gaussian_model.py:347-418) replace a parameter and edit
    state[p]['exp_avg'/'exp_avg_sq'] in place; the fused step must keep working on the edited state."""
    from d2gs_b200.optim import FusedAdam
    dev = cuda_device
    p = torch.nn.Parameter(torch.randn(100, 3, device=dev))
    opt = FusedAdam([{"params": [p], "lr": 0.01, "name": "xyz"}], lr=0.0, eps=1e-15)
    p.grad = torch.randn_like(p); opt.step()
    group = opt.param_groups[0]
    stored = opt.state.get(group["params"][0], None)
    ext = torch.randn(20, 3, device=dev)
    stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(ext)), dim=0)
    stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
    del opt.state[group["params"][0]]
    group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
    opt.state[group["params"][0]] = stored
    q = group["params"][0]
    q.grad = torch.randn_like(q)
    before = q.detach().clone()
    opt.step()
    torch.cuda.synchronize()
    assert q.shape == (120, 3) and not torch.equal(before, q.detach()) and float(opt.state[q]["step"]) == 2.0


def test_densification_stats_kernel(cuda_device):
    from d2gs_b200.optim import add_densification_stats
    dev = cuda_device
    P = 10007
    g = torch.Generator().manual_seed(2)
    vs = torch.zeros(P, 3, device=dev, requires_grad=True)
    vs.grad = torch.randn(P, 3, generator=g).to(dev)
    filt = (torch.rand(P, generator=g) > 0.4).to(dev)
    acc, den = torch.rand(P, 1, generator=g).to(dev), torch.randint(0, 5, (P, 1), generator=g).float().to(dev)
    acc_ref, den_ref = acc.clone(), den.clone()
    acc_ref[filt] += torch.norm(vs.grad[filt, :2], dim=-1, keepdim=True)          # scene/gaussian_model.py:485-486
    den_ref[filt] += 1
    add_densification_stats(acc, den, vs, filt)
    torch.cuda.synchronize()
    assert torch.allclose(acc, acc_ref, rtol=1e-6, atol=1e-7) and torch.equal(den, den_ref)


def test_fused_adam_speed_note(cuda_device):
    """Not an assertion on speed: records the time of one optimiser step over the C3 surfel tables for DESIGN.md."""
    from d2gs_b200.optim import FusedAdam
    dev = cuda_device
    P = 300000
    shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 2), (P, 4), (P, 8)]
    res = {}
    for name, cls, kw in (("torch_foreach", torch.optim.Adam, {}), ("torch_fused", torch.optim.Adam, {"fused": True}), ("d2gs", FusedAdam, {})):
        ps, gs = _groups(dev, 3, shapes)
        opt = cls(gs, lr=0.0, eps=1e-15, **kw)
        for p in ps:
            p.grad = torch.randn_like(p)
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            opt.step()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20.0
    print("ADAM_STEP_MS", res)
    assert res["d2gs"] > 0
