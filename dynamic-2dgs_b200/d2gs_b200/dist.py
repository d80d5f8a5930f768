"""View-sharded data parallelism for the render hot path.

The path shards by camera view: every rank holds a full replica of the surfel parameters and the deformation
network, renders its own views, and the only exchange is the parameter gradient — ONE all-reduce per step over a flat
bucket every ``.grad`` is a view of (no pack/unpack), plus the small densification statistics the trainer keeps
(scene/gaussian_model.py:484-486, train_gui.py:389-391).  One process per GPU, ``torch.distributed`` (NCCL over
NVLink on the B200 box, gloo in the CPU tests).  The reference has no multi-GPU support at all.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def view_for(step: int, rank: int, world: int, n_views: int) -> int:
    """Round-robin view assignment: step s, rank r -> view (s*world + r) mod n_views."""
    return (step * world + rank) % n_views


def views_of_rank(rank: int, world: int, n_views: int) -> List[int]:
    """Static shard of a view list (rank r owns r, r+world, ...)."""
    return list(range(rank, n_views, world))


class FlatGradBucket:
    """All parameter gradients as views of one contiguous fp32 buffer.

    After ``attach()`` autograd accumulates straight into the bucket (``.grad`` is pre-set to a view, so
    AccumulateGrad adds in place); ``all_reduce()`` is then a single collective over the whole buffer."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self) -> None:
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def all_reduce(self, group=None, average: bool = False) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))

    def nbytes(self) -> int:
        return self.numel * 4


def reduce_densification_stats(viewspace_grad_norm: torch.Tensor, visible: torch.Tensor, radii: torch.Tensor, group=None):
    """Keeps densification replica-consistent: SUM of the per-surfel screen-space gradient norms and visibility counts
    (add_densification_stats, scene/gaussian_model.py:484-486) and MAX of the screen radii (train_gui.py:389-391)."""
    counts = visible.to(torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.stack([viewspace_grad_norm.reshape(-1).float() * counts, counts])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        r = radii.clone()
        dist.all_reduce(r, op=dist.ReduceOp.MAX, group=group)
        return packed[0], packed[1], r
    return viewspace_grad_norm.reshape(-1).float() * counts, counts, radii
