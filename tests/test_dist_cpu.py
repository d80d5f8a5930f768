"""CPU, world_size 2 over gloo: the host-side logic of the view-sharded path — round-robin view assignment, the flat
gradient bucket (every .grad a view, one all-reduce) and the densification-stat reductions.  The CUDA kernels are not
involved; a small differentiable stand-in plays the renderer so that  sum over views of single-rank grads ==
all-reduced grads  can be checked exactly (SURVEY.md §8(e) "equivalence test")."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401
from d2gs_b200 import dist as ddist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _model(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(50, 3, generator=g)), torch.nn.Parameter(torch.randn(50, 16, 3, generator=g)),
            torch.nn.Parameter(torch.randn(7, generator=g))]


def _fake_render_loss(params, view):
    a, b, c = params
    t = 0.1 * (view + 1)
    return (torch.sin(a * t).sum() + (b * b * t).mean() + (c * t).pow(3).sum())


def _worker(rank, world, port, n_views, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = _model()
    bucket = ddist.FlatGradBucket(params, direct=bool(out.get("direct", True)))
    seen = []
    for s in range(steps):
        bucket.zero()
        v = ddist.view_for(s, rank, world, n_views)
        seen.append(v)
        _fake_render_loss(params, v).backward()
        bucket.all_reduce()
        lo, hi = bucket.flat.data_ptr(), bucket.flat.data_ptr() + bucket.nbytes()
        assert all(lo <= p.grad.data_ptr() < hi for p in params)      # every .grad is a slice of the bucket
    gn = torch.full((50, 1), float(rank + 1)); vis = torch.arange(50) % (rank + 2) == 0
    radii = torch.arange(50, dtype=torch.int32) * (1 if rank == 0 else -1) + (0 if rank == 0 else 60)
    acc, cnt, rmax = ddist.reduce_densification_stats(gn, vis, radii)
    if rank == 0:
        torch.save({"flat": bucket.flat.clone(), "seen": seen, "acc": acc, "cnt": cnt, "rmax": rmax}, out["path"])
    dist.destroy_process_group()


def test_view_sharding_is_a_partition():
    for world in (1, 2, 4, 8):
        for n_views in (100, 7):
            steps = -(-n_views // world)
            got = sorted(ddist.view_for(s, r, world, n_views) for s in range(steps) for r in range(world))[:n_views]
            assert set(got) == set(range(n_views)) or n_views % world
            shards = [ddist.views_of_rank(r, world, n_views) for r in range(world)]
            assert sorted(sum(shards, [])) == list(range(n_views))


@pytest.mark.parametrize("direct", [True, False])
def test_flat_bucket_allreduce_equals_sum_over_views(tmp_path, direct):
    world, n_views, steps = 2, 6, 3
    path = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, steps, {"path": path, "direct": direct}), nprocs=world, join=True)
    got = torch.load(path)
    # single-process reference: gradient of the LAST step's views summed over ranks
    params = _model()
    last_views = [ddist.view_for(steps - 1, r, world, n_views) for r in range(world)]
    total = sum(_fake_render_loss(params, v) for v in last_views)
    total.backward()
    pad = lambda g: torch.cat([g.reshape(-1), torch.zeros(-g.numel() % ddist.FlatGradBucket.ALIGN)])      # slots are 256-B aligned
    ref = torch.cat([pad(p.grad) for p in params])
    assert torch.allclose(got["flat"], ref, rtol=1e-6, atol=1e-6)
    assert got["seen"] == [ddist.view_for(s, 0, world, n_views) for s in range(steps)]
    # densification statistics: SUM of masked norms / counts, MAX of radii
    vis0, vis1 = (torch.arange(50) % 2 == 0).float(), (torch.arange(50) % 3 == 0).float()
    assert torch.allclose(got["cnt"], vis0 + vis1) and torch.allclose(got["acc"], 1.0 * vis0 + 2.0 * vis1)
    assert torch.equal(got["rmax"], torch.maximum(torch.arange(50, dtype=torch.int32), 60 - torch.arange(50, dtype=torch.int32)))


def test_bucket_single_process_noop():
    params = _model()
    b = ddist.FlatGradBucket(params, direct=False)
    _fake_render_loss(params, 2).backward()
    before = b.flat.clone()
    b.all_reduce()          # no process group: must be a no-op
    assert torch.equal(before, b.flat) and b.nbytes() >= 4 * sum(p.numel() for p in params)
    b.zero()
    assert not any(p.grad.any() for p in params)


class _ClaimingSquare(torch.autograd.Function):
    """Stand-in for the CUDA backward wrappers: writes the parameter gradient into the slot the bucket hands out."""

    @staticmethod
    def forward(ctx, w):
        ctx.save_for_backward(w.detach())
        return (w * w).sum()

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        out = ddist.claim(w, zeroed=False)
        if out is None:
            out = torch.empty_like(w)
        torch.mul(w, 2.0 * g, out=out)
        return out


def test_direct_bucket_adopts_kernel_written_gradients():
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.randn(40, 3))
    small = torch.nn.Parameter(torch.randn(5))
    unused = torch.nn.Parameter(torch.randn(4))
    b = ddist.FlatGradBucket([big, small, unused], large_numel=100)
    assert b.params[0] is small and b.params[-1] is big            # small slots first, large ones last
    try:
        for step in range(3):
            b.zero()
            assert big.grad is None and small.grad is None
            loss = _ClaimingSquare.apply(big) + _ClaimingSquare.apply(small)
            if step == 1:
                loss = loss + _ClaimingSquare.apply(small) + big.sum()      # parameters used twice: autograd sums out of place
            if step == 2:
                loss = _ClaimingSquare.apply(small)                          # the large slot is not written this step
            loss.backward()
            if step == 0:      # autograd adopted the views: no accumulate, no copy
                assert big.grad.data_ptr() == b.slots[-1].view.data_ptr() and small.grad.data_ptr() == b.slots[0].view.data_ptr()
            b.all_reduce()
            lo, hi = b.flat.data_ptr(), b.flat.data_ptr() + b.nbytes()
            assert all(lo <= p.grad.data_ptr() < hi for p in (big, small, unused))
            eb = {0: 2 * big.data, 1: 2 * big.data + 1, 2: torch.zeros_like(big)}[step]
            es = {0: 2 * small.data, 1: 4 * small.data, 2: 2 * small.data}[step]
            assert torch.allclose(big.grad, eb) and torch.allclose(small.grad, es) and not unused.grad.any()
            assert all(s_.view.data_ptr() % 256 == b.flat.data_ptr() % 256 for s_ in b.slots)
            assert float(b.flat.sum()) == pytest.approx(float(small.grad.sum() + big.grad.sum()), rel=1e-5, abs=1e-5)
    finally:
        b.detach()


def _early_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.randn(50, 3))       # "surfel table": final after the first backward stage
    late = torch.nn.Parameter(torch.randn(7))            # gets its gradient later in the same backward
    b = ddist.FlatGradBucket([late, table], large_numel=1 << 20, early=[table])
    assert b.params[-1] is table and b.early_begin == b.numel - 192      # early slots are laid out last (50*3 -> 3 x 64 floats)

    class Notify(torch.autograd.Function):            # stands for the end of the rasterizer backward
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            ddist.grads_ready("raster")
            return g

    launched = []
    for step in range(3):
        b.zero()
        scale = float(rank + 1 + step)
        y = _ClaimingSquare.apply(table)                  # scalar sum(table^2); its backward claims the table's slot
        z = Notify.apply(late * 1.0)
        (scale * y).backward()                            # stage 1: the table gradient is written into its slot
        (scale * (z * z).sum()).backward()                # stage 2: Notify.backward -> early all-reduce, then late's gradient
        launched.append(b._early_work is not None)
        b.all_reduce()
        exp_table = sum(2.0 * (r + 1 + step) for r in range(world)) * table.data
        exp_late = sum(2.0 * (r + 1 + step) for r in range(world)) * late.data
        assert torch.allclose(table.grad, exp_table, rtol=1e-5), step
        assert torch.allclose(late.grad, exp_late, rtol=1e-5), step
    if rank == 0:
        torch.save({"launched": launched}, out["path"])
    dist.destroy_process_group()


def test_early_allreduce_of_declared_tables(tmp_path):
    world = 2
    path = str(tmp_path / "early.pt")
    mp.spawn(_early_worker, args=(world, _free_port(), {"path": path}), nprocs=world, join=True)
    assert torch.load(path)["launched"] == [True, True, True]      # the early collective really ran ahead of the final one


class _TwoConsumers(torch.autograd.Function):
    """Stand-in for _RasterizeSurfelsRaw.backward (raster.py): the gradient of a table is ALSO the gradient of a delta
    that an earlier stage of the graph produced; the table's slot is reduced early, so the delta must not alias it."""

    @staticmethod
    def forward(ctx, table, delta):
        ctx.save_for_backward(table.detach())
        ctx.delta = delta.detach()
        return ((table + delta) ** 2).sum()

    @staticmethod
    def backward(ctx, g):
        (table,) = ctx.saved_tensors
        out = ddist.claim(table, zeroed=False)
        torch.mul(table + ctx.delta, 2.0 * g, out=out)
        g_delta = out.clone() if ddist.reduces_early(table) else out.view_as(out)
        ddist.grads_ready("raster")                     # launches the in-place early all-reduce of `out`
        return out, g_delta


def _alias_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    table = torch.nn.Parameter(torch.randn(50, 3))
    net = torch.nn.Parameter(torch.randn(3))              # "deformation network": consumes the delta's gradient downstream
    b = ddist.FlatGradBucket([net, table], early=[table])
    assert ddist.reduces_early(table) and not ddist.reduces_early(net)
    b.zero()
    delta = net[None, :].expand(50, 3) * 0.5              # d(delta)/d(net) consumes the LOCAL gradient of the table, downstream
    scale = float(rank + 1)
    (scale * _TwoConsumers.apply(table, delta)).backward()
    b.all_reduce()
    # the downstream consumer must have seen the LOCAL table gradient: the all-reduced net gradient is the sum of the
    # per-rank contributions computed independently (were it fed from the slot while it is being reduced, rank r could
    # see the other ranks' values as well and the total would come out too large)
    tot = sum(2.0 * (r + 1) for r in range(world))
    want_table = tot * (table.data + delta.detach())
    assert torch.allclose(table.grad, want_table, rtol=1e-5)
    assert torch.allclose(net.grad, 0.5 * want_table.sum(0), rtol=1e-5)
    if rank == 0:
        torch.save({"ok": True}, out["path"])
    dist.destroy_process_group()


def test_early_reduced_slot_is_never_aliased_downstream(tmp_path):
    """with N > 1 the gradient handed to the deformation deltas must be a private copy, taken before the
    early all-reduce starts summing other ranks' values into the parameter's bucket slot."""
    params = _model()
    b = ddist.FlatGradBucket(params, early=[params[1]])
    try:
        assert not ddist.reduces_early(params[1])          # no process group: nothing is reduced, aliasing is safe
    finally:
        b.detach()
    world = 2
    path = str(tmp_path / "alias.pt")
    mp.spawn(_alias_worker, args=(world, _free_port(), {"path": path}), nprocs=world, join=True)
    assert torch.load(path)["ok"]


def _staged_worker(rank, world, port, out):
    """Two named stages: 'raster' reports first, 'deform' second, the head goes with the final collective."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    head = torch.nn.Parameter(torch.randn(7))
    mid = torch.nn.Parameter(torch.randn(40, 2))
    tab = torch.nn.Parameter(torch.randn(50, 3))
    b = ddist.FlatGradBucket([head, mid, tab], stages={"raster": [tab], "deform": [mid]})
    assert b.params == [head, mid, tab] and set(b.stage_ranges) == {"raster", "deform"}
    assert b.stage_ranges["deform"][1] == b.stage_ranges["raster"][0] and b.early_begin == b.stage_ranges["deform"][0]
    res = []
    for step, report in enumerate([("raster", "deform"), ("raster",), ()]):
        b.zero()
        for p_, scale in ((head, 1.0), (mid, 2.0), (tab, 3.0)):
            g = ddist.claim(p_, zeroed=False)
            g.copy_(torch.full_like(p_, scale * (rank + 1) + step))
            p_.grad = g
        for st in report:
            ddist.grads_ready(st)
        launched = sorted(b._stage_work)
        b.all_reduce()
        res.append((launched, float(head.grad[0]), float(mid.grad[0, 0]), float(tab.grad[0, 0])))
    if rank == 0:
        torch.save(res, out["path"])
    dist.destroy_process_group()


def test_staged_allreduce_with_two_stages(tmp_path):
    world = 2
    path = str(tmp_path / "staged.pt")
    mp.spawn(_staged_worker, args=(world, _free_port(), {"path": path}), nprocs=world, join=True)
    res = torch.load(path)
    assert [r[0] for r in res] == [["deform", "raster"], ["raster"], []]
    for step, (_, h, m, t) in enumerate(res):       # sum over ranks 1 and 2 of scale * (rank + 1) + step
        assert (h, m, t) == (1.0 * 3 + 2 * step, 2.0 * 3 + 2 * step, 3.0 * 3 + 2 * step)


def test_balanced_view_schedule_is_an_epoch_of_equal_cost_groups():
    import random
    rnd = random.Random(0)
    costs = [rnd.uniform(1.0, 2.0) for _ in range(96)]
    sched = ddist.balanced_view_schedule(costs, 8)
    assert len(sched) == 12 and all(len(g) == 8 for g in sched)
    assert sorted(v for g in sched for v in g) == list(range(96))                    # every view once per epoch
    spread = max(max(costs[v] for v in g) - min(costs[v] for v in g) for g in sched)
    assert spread < 0.2 * (max(costs) - min(costs))                                   # the views of a step cost about the same
    naive = max(max(costs[v] for v in range(s * 8, s * 8 + 8)) - min(costs[v] for v in range(s * 8, s * 8 + 8)) for s in range(12))
    assert spread < 0.5 * naive
    # 100 views on 8 ranks: the last group wraps around; still 8 per step
    s100 = ddist.balanced_view_schedule([rnd.random() for _ in range(100)], 8)
    assert len(s100) == 13 and all(len(g) == 8 for g in s100) and set(v for g in s100 for v in g) == set(range(100))
