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
    bucket = ddist.FlatGradBucket(params)
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in params)
    seen = []
    for s in range(steps):
        bucket.zero()
        v = ddist.view_for(s, rank, world, n_views)
        seen.append(v)
        _fake_render_loss(params, v).backward()
        assert all(p.grad.data_ptr() == q for p, q in zip(params, out["ptrs"])) if out.get("ptrs") else True
        bucket.all_reduce()
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


def test_flat_bucket_allreduce_equals_sum_over_views(tmp_path):
    world, n_views, steps = 2, 6, 3
    path = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, steps, {"path": path}), nprocs=world, join=True)
    got = torch.load(path)
    # single-process reference: gradient of the LAST step's views summed over ranks
    params = _model()
    last_views = [ddist.view_for(steps - 1, r, world, n_views) for r in range(world)]
    total = sum(_fake_render_loss(params, v) for v in last_views)
    total.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in params])
    assert torch.allclose(got["flat"], ref, rtol=1e-6, atol=1e-6)
    assert got["seen"] == [ddist.view_for(s, 0, world, n_views) for s in range(steps)]
    # densification statistics: SUM of masked norms / counts, MAX of radii
    vis0, vis1 = (torch.arange(50) % 2 == 0).float(), (torch.arange(50) % 3 == 0).float()
    assert torch.allclose(got["cnt"], vis0 + vis1) and torch.allclose(got["acc"], 1.0 * vis0 + 2.0 * vis1)
    assert torch.equal(got["rmax"], torch.maximum(torch.arange(50, dtype=torch.int32), 60 - torch.arange(50, dtype=torch.int32)))


def test_bucket_single_process_noop():
    params = _model()
    b = ddist.FlatGradBucket(params)
    _fake_render_loss(params, 2).backward()
    before = b.flat.clone()
    b.all_reduce()          # no process group: must be a no-op
    assert torch.equal(before, b.flat) and b.nbytes() == 4 * sum(p.numel() for p in params)
    b.zero()
    assert not any(p.grad.any() for p in params)
