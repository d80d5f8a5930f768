"""Spatially sorted storage of the surfel tables.

The deformation kernels process surfels in a Morton order of their centres so that the 32 surfels of a warp share
their control nodes (``deform.processing_order``).  When the tables are STORED in arbitrary order that processing order
turns every per-surfel access into a scattered 12-32-byte gather (``deform_fwd``: 50 MB of DRAM traffic for 35 MB of
operands, ``profiles/ncu_r1i.md``).  ``permute_surfels_`` re-orders every per-surfel tensor of a model — parameters,
optimiser moments, densification statistics — with one permutation, in place, so the processing order becomes the
identity and those gathers become coalesced streams; the per-tile gather of the blend kernels also gets neighbouring
records from neighbouring addresses.

Where to call it in the reference trainer: right after densification / pruning
(``train_gui.py:413-423`` -> ``gaussians.densify_and_prune``), which already rebuilds every per-surfel tensor and the
optimiser state — once every 100 iterations, ~0.3 ms at 300 k surfels:

    from d2gs_b200.layout import morton_permutation, permute_surfels_
    permute_surfels_(gaussians, morton_permutation(gaussians.get_xyz), optimizers=[gaussians.optimizer])

Nothing depends on the storage order: results are identical up to the order of the floating-point gradient sums."""
from __future__ import annotations

from typing import Iterable

import torch


def morton_permutation(xyz: torch.Tensor) -> torch.Tensor:
    """int64 permutation that sorts the centres along a 30-bit Morton curve (libd2gs.so: d2gs_deform_order)."""
    from .deform import processing_order
    return processing_order(xyz.detach()).long()


def _permute_tensor(t: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    return t.index_select(0, perm).contiguous()


def permute_surfels_(model, perm: torch.Tensor, optimizers: Iterable[torch.optim.Optimizer] = ()) -> int:
    """Re-orders, IN PLACE, every tensor of ``model`` whose leading dimension is the surfel count (``nn.Parameter``s keep
    their identity, so optimiser state keys and autograd bookkeeping stay valid), the per-parameter optimiser state
    (``exp_avg`` / ``exp_avg_sq`` and any other state tensor of that shape) and plain per-surfel buffers
    (``xyz_gradient_accum``, ``denom``, ``max_radii2D``, ...).  Returns the number of tensors touched."""
    P = int(perm.numel())
    perm = perm.to(torch.long)
    if P == 0 or int(perm.min()) < 0 or int(perm.max()) >= P or not bool((torch.bincount(perm, minlength=P) == 1).all()):
        raise ValueError("perm must be a permutation of 0..P-1")
    touched, seen = 0, set()
    params = []
    with torch.no_grad():
        names = list(getattr(model, "_parameters", {}).items()) + list(getattr(model, "_buffers", {}).items()) + list(vars(model).items())
        for name, t in names:
            if not torch.is_tensor(t) or t.dim() == 0 or t.shape[0] != P or id(t) in seen:
                continue
            seen.add(id(t))
            p = perm.to(t.device)
            if isinstance(t, torch.nn.Parameter):
                t.data = _permute_tensor(t.data, p)
                if t.grad is not None:
                    t.grad = None          # a gradient of the old order must not survive
                params.append(t)
            else:
                new = _permute_tensor(t, p)
                if name in getattr(model, "_buffers", {}):
                    model._buffers[name] = new
                else:
                    setattr(model, name, new)
            touched += 1
        for opt in optimizers:
            for prm in params:
                st = opt.state.get(prm)
                if not st:
                    continue
                for k, v in list(st.items()):
                    if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == P:
                        st[k] = _permute_tensor(v, perm.to(v.device))
                        touched += 1
    # cached processing orders refer to the old storage order
    for obj in (model, getattr(model, "deform", None)):
        if obj is not None and getattr(obj, "_order_cache", None) is not None:
            object.__setattr__(obj, "_order_cache", None)
    return touched
