"""View-sharded data parallelism for the render hot path.

The path shards by camera view: every rank holds a full replica of the surfel parameters and the deformation
network, renders its own views, and the only exchange is the parameter gradient — ONE all-reduce per step over a flat
bucket the backward kernels write into directly (no AccumulateGrad kernels, no pack/unpack), plus the small densification statistics the trainer keeps
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


def balanced_view_schedule(costs, world: int) -> List[List[int]]:
    """Cost-balanced view schedule for synchronous data parallelism: every step waits for its slowest rank, so the ``world``
    views of a step should cost the same.  Views are sorted by ``costs`` (e.g. the (surfel, tile) instance count of a view
    from an earlier epoch) and cut into consecutive groups of ``world``; the group order is interleaved (cheap, expensive,
    cheap, ...) so that the epoch has no trend.  Every view appears exactly once per epoch when ``len(costs)`` is a multiple
    of ``world`` (the remainder wraps around to the cheapest views).  Returns ``schedule[step % len(schedule)][rank]``."""
    order = sorted(range(len(costs)), key=lambda v: (float(costs[v]), v))
    groups = [[order[(g * world + r) % len(order)] for r in range(world)] for g in range((len(order) + world - 1) // world)]
    lo, hi, out = 0, len(groups) - 1, []
    while lo <= hi:
        out.append(groups[lo]); lo += 1
        if lo <= hi:
            out.append(groups[hi]); hi -= 1
    return out


def views_of_rank(rank: int, world: int, n_views: int) -> List[int]:
    """Static shard of a view list (rank r owns r, r+world, ...)."""
    return list(range(rank, n_views, world))


class _Slot:
    __slots__ = ("view", "claimed", "dirty", "large", "early_group")

    def __init__(self, view, large):
        self.view, self.claimed, self.dirty, self.large = view, False, False, large
        self.early_group = None      # ("default" | process group) when the slot is part of an early all-reduce


_SLOTS = {}   # data_ptr of a parameter -> _Slot of the bucket that owns it


def claim(t: Optional[torch.Tensor], zeroed: bool) -> Optional[torch.Tensor]:
    """Called by the backward wrappers (raster.py, deform.py) with the tensor they saved for a parameter: returns that
    parameter's slice of the gradient bucket, shaped like ``t``, so the CUDA backward writes the gradient where the
    all-reduce will read it (autograd then adopts the returned view as ``.grad`` without an accumulate kernel).
    ``zeroed``: the kernel accumulates with atomics and needs zeros.  Returns None when no bucket owns the tensor or
    the slot was already handed out this step (a parameter used twice: autograd sums out of place and
    ``FlatGradBucket.finalize`` copies the sum back)."""
    if t is None or not _SLOTS:
        return None
    s = _SLOTS.get(t.data_ptr())
    if s is None or s.claimed or s.view.numel() != t.numel() or not t.is_contiguous() or s.view.device != t.device:
        return None
    if zeroed and s.dirty:
        s.view.zero_()
    s.claimed, s.dirty = True, True
    return s.view.view(t.shape)


def reduces_early(t: Optional[torch.Tensor]) -> bool:
    """True when the bucket slot of parameter ``t`` will be summed across ranks by the EARLY all-reduce that
    ``grads_ready("raster")`` launches (in place, asynchronously).  From that moment the slot must not be read by the
    rest of the backward pass: a wrapper that hands the same gradient to a second consumer (the deformation deltas of
    ``_xyz`` / ``_rotation``) gives that consumer a private copy instead of an alias."""
    if t is None or not _SLOTS:
        return False
    s = _SLOTS.get(t.data_ptr())
    if s is None or s.early_group is None:
        return False
    if not (dist.is_available() and dist.is_initialized()):
        return False
    return dist.get_world_size(None if s.early_group == "default" else s.early_group) > 1


_ACTIVE = []  # weak references to the direct-mode buckets, for grads_ready()


def grads_ready(stage: str) -> None:
    """Called by the backward wrappers when a stage of the backward pass has written all its parameter gradients
    ("raster": the per-surfel tables, after the per-surfel rasterizer backward; "deform": the hyper-coordinate table and the
    node geometry, after the deformation blend backward).  Buckets that declared parameters for that stage start its
    all-reduce now, on the collective's own stream, overlapped with the rest of the backward."""
    for ref in list(_ACTIVE):
        b = ref()
        if b is None:
            _ACTIVE.remove(ref)
        else:
            b._on_stage(stage)


class FlatGradBucket:
    """All parameter gradients as slices of one contiguous fp32 buffer, reduced with ONE collective (two when part of
    it can start early).

    ``direct=True`` (default): ``.grad`` starts every step as None and the backward kernels write into the bucket
    through ``claim()``; autograd adopts those views, so a step costs no AccumulateGrad kernels and no pack.
    Small slots (< ``large_numel`` elements) sit first and are cleared by one fill per step; large slots (the per-surfel
    tables) are fully overwritten by the rasterizer backward and only cleared when a step does not claim them.
    ``early``: parameters whose gradient is final as soon as the rasterizer backward has run (the surfel tables; NOT
    parameters that also feed the deformation).  They are laid out last and contiguously, and their all-reduce is
    issued from ``grads_ready("raster")`` while the deformation / MLP backward still runs.
    ``direct=False``: ``.grad`` is pre-set to the views and autograd accumulates in place."""

    ALIGN = 64   # floats

    def __init__(self, params: Iterable[torch.nn.Parameter], direct: bool = True, large_numel: int = 1 << 20,
                 early: Optional[Iterable[torch.nn.Parameter]] = None, stages=None):
        """``stages``: ``{"raster": [...], "deform": [...]}`` — parameters whose gradient is final when the named stage of
        the backward pass reports ``grads_ready(stage)``; ``early=[...]`` is shorthand for ``{"raster": [...]}``."""
        ps = [p for p in params if p.requires_grad]
        assert ps, "no trainable parameters"
        stage_of = {}
        if direct:
            for name, plist in ({"raster": early or []} if stages is None else dict(stages)).items():
                for p in plist:
                    if p.requires_grad:
                        stage_of[id(p)] = name
        is_early = lambda p: id(p) in stage_of
        is_large = lambda p: p.numel() >= large_numel or is_early(p)
        stage_names = []
        for p in ps:
            if is_early(p) and stage_of[id(p)] not in stage_names:
                stage_names.append(stage_of[id(p)])
        # staged parameters sit last, one contiguous range per stage; the stage that reports first ("raster") is the last range
        stage_names.sort(key=lambda n: 0 if n == "raster" else -1)
        self.params = [p for p in ps if not is_large(p)] + [p for p in ps if is_large(p) and not is_early(p)]
        for name in stage_names:
            self.params += [p for p in ps if is_early(p) and stage_of[id(p)] == name]
        dev = self.params[0].device
        # every slot starts on a 256-byte boundary, like a tensor from the CUDA caching allocator: the backward kernels
        # store gradients with 16-byte vector instructions
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = sum(pad(p.numel()) for p in self.params)
        self.n_small = sum(pad(p.numel()) for p in self.params if not is_large(p))
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.direct = direct
        self.slots: List[_Slot] = []
        self.early_slots: List[_Slot] = []
        self.stage_ranges = {}        # name -> [begin, end, slots]
        off = 0
        self.early_begin = None
        for p in self.params:
            sl = _Slot(self.flat[off: off + p.numel()].view_as(p), is_large(p))
            self.slots.append(sl)
            if is_early(p):
                if self.early_begin is None:
                    self.early_begin = off
                sl.early_group = "default"
                self.early_slots.append(sl)
                r = self.stage_ranges.setdefault(stage_of[id(p)], [off, off, []])
                r[1] = off + pad(p.numel())
                r[2].append(sl)
            off += pad(p.numel())
        self._stage_work = {}         # name -> async work handle of this step
        self._group = None
        if direct:
            import weakref
            _ACTIVE.append(weakref.ref(self))
        self.attach()

    @property
    def _early_work(self):
        """The "raster" stage's work handle (kept for callers and tests written against the single early stage)."""
        return self._stage_work.get("raster")

    def attach(self) -> None:
        for p, s in zip(self.params, self.slots):
            if self.direct:
                _SLOTS[p.data_ptr()] = s
                p.grad = None
            else:
                p.grad = s.view

    def detach(self) -> None:
        for p in self.params:
            _SLOTS.pop(p.data_ptr(), None)
        _ACTIVE[:] = [r for r in _ACTIVE if r() is not None and r() is not self]

    def zero(self) -> None:
        """Start of a step."""
        if not self.direct:
            self.flat.zero_()
            return
        for w in self._stage_work.values():    # a step that never reached all_reduce()
            w.wait()
        self._stage_work = {}
        if self.n_small:
            self.flat[: self.n_small].zero_()
        for p, s in zip(self.params, self.slots):
            p.grad = None
            s.claimed = False
            if not s.large:
                s.dirty = False

    begin_step = zero

    def set_group(self, group) -> None:
        """Process group used by the early all-reduce (default group when never called)."""
        self._group = group
        for sl in self.early_slots:
            sl.early_group = "default" if group is None else group

    def _on_stage(self, stage: str) -> None:
        r = self.stage_ranges.get(stage)
        if r is None or stage in self._stage_work:
            return
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1):
            return
        if not all(s.claimed for s in r[2]):
            return            # some table was not written by the kernels this step: it goes with the final all-reduce
        self._stage_work[stage] = dist.all_reduce(self.flat[r[0]:r[1]], op=dist.ReduceOp.SUM, group=self._group, async_op=True)

    def finalize(self) -> None:
        """After backward: every ``.grad`` is its bucket slice and the slice holds this step's gradient."""
        if not self.direct:
            return
        early = {id(sl) for name in self._stage_work for sl in self.stage_ranges[name][2]}
        for p, s in zip(self.params, self.slots):
            g = p.grad
            if g is None:
                if id(s) in early:
                    raise RuntimeError("a parameter declared `early` was claimed but received no gradient")
                if s.dirty:                # stale data of an earlier step, or handed out but never returned to autograd
                    s.view.zero_()
                    s.dirty = False
                p.grad = s.view
            elif g.data_ptr() != s.view.data_ptr():
                if id(s) in early:
                    raise RuntimeError("a parameter declared `early` received a second gradient after the rasterizer backward")
                s.view.copy_(g)
                s.dirty = True
                p.grad = s.view

    def all_reduce(self, group=None, average: bool = False) -> None:
        self.finalize()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self._stage_work:
                # what no stage has started yet: the unstaged head plus the ranges of stages that never reported
                end = self.early_begin
                pending = [(b, e) for name, (b, e, _) in self.stage_ranges.items() if name not in self._stage_work]
                for b, e in sorted(pending):
                    if b == end:
                        end = e               # contiguous with the head: one collective
                if end > 0:
                    dist.all_reduce(self.flat[:end], op=dist.ReduceOp.SUM, group=group)
                for b, e in sorted(pending):
                    if b >= end:
                        dist.all_reduce(self.flat[b:e], op=dist.ReduceOp.SUM, group=group)
                for w in self._stage_work.values():
                    w.wait()
                self._stage_work = {}
            else:
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
