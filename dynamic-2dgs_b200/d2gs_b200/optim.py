"""Fused optimiser step (libd2gs.so: d2gs_adam_step) and densification statistics.

``FusedAdam`` is a drop-in for the two optimisers the reference builds with ``torch.optim.Adam(l, lr=0.0, eps=1e-15)``
(scene/gaussian_model.py:181-203, scene/deform_model.py) and steps in train_gui.py:426-432: same ``param_groups``, same
``state[p] = {"step", "exp_avg", "exp_avg_sq"}`` layout — the reference's densification code edits exactly those
entries (cat_tensors_to_optimizer / _prune_optimizer / replace_tensor_to_optimizer) and keeps working — same update
rule (torch's ``_single_tensor_adam``: no weight decay, no amsgrad), but ONE kernel launch per ``step()``.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise NotImplementedError("FusedAdam implements the reference's configuration: no weight decay, no amsgrad")
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        per_device = {}
        keep = []           # contiguous gradient copies must outlive the launch
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr, eps = float(group["lr"]), float(group["eps"])
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam updates CUDA parameters only (no CPU path)")
                if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam needs contiguous float32 parameters and dense gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                t = float(st["step"])
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                    keep.append(g)
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if m.shape != p.shape or v.shape != p.shape or not m.is_contiguous() or not v.is_contiguous():
                    raise RuntimeError("optimizer state does not match its parameter (stale state after a parameter was replaced?)")
                d = _lib.AdamTensor()
                d.param, d.grad, d.exp_avg, d.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                d.numel = p.numel()
                d.beta1, d.beta2, d.eps = beta1, beta2, eps
                d.step_size = lr / (1.0 - beta1 ** t)                       # python floats = double, like torch
                d.bias_correction2_sqrt = math.sqrt(1.0 - beta2 ** t)
                per_device.setdefault(p.device, []).append(d)
        for dev, descs in per_device.items():
            arr = (_lib.AdamTensor * len(descs))(*descs)
            with torch.cuda.device(dev):
                _lib.check(L.d2gs_adam_step(arr, len(descs), _stream(dev)), "d2gs_adam_step")
        return loss


def add_densification_stats(xyz_gradient_accum: torch.Tensor, denom: torch.Tensor, viewspace_point_tensor: torch.Tensor,
                            update_filter: torch.Tensor) -> None:
    """scene/gaussian_model.py:484-486 as one kernel:  accum[f] += |viewspace.grad[f, :2]|,  denom[f] += 1."""
    g = viewspace_point_tensor.grad if viewspace_point_tensor.grad is not None else viewspace_point_tensor
    if not g.is_cuda:
        raise RuntimeError("add_densification_stats needs CUDA tensors")
    g = g.detach().float().contiguous()
    f = update_filter.to(torch.bool).contiguous()
    P = int(g.shape[0])
    assert xyz_gradient_accum.is_contiguous() and denom.is_contiguous() and xyz_gradient_accum.numel() == P and denom.numel() == P
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().d2gs_densification_stats(P, g.data_ptr(), int(g.shape[1]), f.data_ptr(), xyz_gradient_accum.data_ptr(),
                                                       denom.data_ptr(), _stream(g.device)), "d2gs_densification_stats")
