"""Timing (not a pytest) of the depth/alpha 3-D Gaussian rasterizer drop-in against the unmodified reference extension:
    python tests/gpu_gs3d_time.py [P ...]
fwd+bwd through the public module API on a seeded cloud at 800x800; CUDA events around `reps` iterations after warm-up.
Modes: reference | ours, global sort + count readback (round 1) | ours, per-tile binning + readback | ours, deferred count |
ours, deferred count replayed as a CUDA graph."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dynamic-2dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np
import torch
from make_gs3d_golden import load_reference_dgr
from d2gs_b200 import _lib, raster, synthetic as syn
import diff_gaussian_rasterization as ours

dev = torch.device("cuda:0")


def scene(P, W=800, H=800, seed=77):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(P, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = (d * rng.uniform(size=(P, 1)) ** (1 / 3)).astype(np.float32)
    s_med = 0.006 * (100_000 / P) ** (1 / 3)
    scales = np.exp(np.log(s_med) + 0.5 * rng.normal(size=(P, 3))).astype(np.float32)
    q = rng.normal(size=(P, 4)); q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    cam = syn.fibonacci_cameras(100, W, H)[37]
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    leaves = dict(means3D=T(xyz), opacities=T(rng.uniform(0.05, 0.95, size=(P, 1))), scales=T(scales), rotations=T(q),
                  shs=T(np.concatenate([rng.normal(size=(P, 1, 3)), 0.1 * rng.normal(size=(P, 15, 3))], 1)))
    for v in leaves.values():
        v.requires_grad_(True)
    cams = dict(tanfovx=float(cam.tanfovx), tanfovy=float(cam.tanfovy), bg=T([0.3, 0.1, 0.7]), viewmatrix=T(cam.world_view_transform),
                projmatrix=T(cam.full_proj_transform), campos=T(cam.camera_center))
    g = (T(rng.normal(size=(3, H, W))), T(rng.normal(size=(1, H, W))), T(rng.normal(size=(1, H, W))))
    return leaves, cams, g, W, H


def make_step(mod, leaves, cams, g, W, H):
    rs = mod.GaussianRasterizationSettings(image_height=H, image_width=W, scale_modifier=1.0, sh_degree=3, prefiltered=False, debug=False, **cams)
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    params = list(leaves.values()) + [m2d]

    def step():
        color, radii, depth, alpha = mod.GaussianRasterizer(rs)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                                shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        return torch.autograd.grad((color * g[0]).sum() + (depth * g[1]).sum() + (alpha * g[2]).sum(), params)
    return step


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for P in [int(a) for a in sys.argv[1:]] or [100_000, 300_000]:
    leaves, cams, g, W, H = scene(P)
    out = {"P": P}
    ref = load_reference_dgr()
    if ref is not None:
        out["reference_ms"] = timed(make_step(ref, leaves, cams, g, W, H))
    step = make_step(ours, leaves, cams, g, W, H)
    _lib.set_option("tile_sort", 0); raster.set_deferred_count(False)
    out["ours_global_sort_sync_ms"] = timed(step)
    _lib.set_option("tile_sort", 1)
    out["ours_tile_binning_sync_ms"] = timed(step)
    raster.set_deferred_count(True, warmup=2, margin=1.5)
    out["ours_deferred_ms"] = timed(step)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        step(); torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            step()
        out["ours_deferred_graph_ms"] = timed(graph.replay)
    raster.set_deferred_count(False)
    print("GS3D_TIME " + json.dumps(out), flush=True)
