"""GPU: the depth/alpha 3-D Gaussian rasterizer drop-in (``diff_gaussian_rasterization`` -> libd2gs.so: d2gs_gs3d_*) against
the CPU oracle (oracle/gs3d_oracle.py), the golden outputs of the unmodified reference extension, and — when
oracle/_ref/diff_gaussian_rasterization travelled to the box — the live reference extension."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_gs3d_golden import gs3d_cases, load_reference_dgr, run_module  # noqa: E402
from oracle import gs3d_oracle as go  # noqa: E402
import util  # noqa: E402

pytestmark = pytest.mark.gpu
CASES = gs3d_cases()
RTOL_MAPS = 1e-4        # north_star: RGB / depth maps within 1e-4 relative fp32
RTOL_GRADS = 5e-4       # norm-wise; gradients sum 10^3-10^4 fp32 terms in nondeterministic order
GRAD_KEYS = ("g_means2D", "g_means3D", "g_opacities", "g_shs", "g_colors_precomp", "g_scales", "g_rotations", "g_cov3D_precomp")


def _golden(name):
    p = os.path.join(HERE, "golden", f"gs3d_golden_{name}.npz")
    return np.load(p) if os.path.exists(p) else None


def _ours(name, device):
    import diff_gaussian_rasterization as ours
    from d2gs_b200 import gs3d
    res = run_module(ours, CASES[name], device)
    st = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in gs3d.export_state().items()}
    return res, st


@pytest.mark.parametrize("name", list(CASES.keys()))
def test_matches_oracle(name, cuda_device):
    res, st = _ours(name, cuda_device)
    c = CASES[name]
    o32 = go.render(dtype=torch.float32, **c["inputs"])
    o64, g64 = go.render_with_grads(c["inputs"], c["g_color"], c["g_depth"], c["g_alpha"], dtype=torch.float64)
    # integer stages against the float32 oracle: a radius sits on a ceil() boundary for ~1 Gaussian in 10^4 (fma contraction)
    rad_o = o32["radii"].numpy()
    assert (res["radii"] == rad_o).mean() >= 0.998
    same = res["radii"] == rad_o
    assert np.array_equal(st["tiles_touched"][same], o32["tiles_touched"].numpy()[same])
    assert abs(st["num_rendered"] - o32["num_rendered"]) <= 0.002 * o32["num_rendered"] + 8
    for k in ("color", "depth", "alpha"):
        assert util.rel_err(res[k], o64[k].numpy()) < RTOL_MAPS, (k, util.rel_err(res[k], o64[k].numpy()))
    assert (st["n_contrib"] == o64["n_contrib"].numpy()).mean() > 0.999
    names = dict(g_means2D="means2D", g_means3D="means3D", g_opacities="opacities", g_shs="shs", g_colors_precomp="colors_precomp",
                 g_scales="scales", g_rotations="rotations", g_cov3D_precomp="cov3D_precomp")
    errs = {k: util.rel_err(res[k], g64[names[k]].numpy().reshape(res[k].shape)) for k in GRAD_KEYS if k in res}
    print(name, errs)
    # G1 (camera inside the cloud, screen-filling Gaussians): fp32 gradients of BOTH implementations sit ~6e-4 from float64
    tol = 2e-3 if name == "G1" else RTOL_GRADS
    assert all(v < tol for v in errs.values()), errs


@pytest.mark.parametrize("name", list(CASES.keys()))
def test_matches_reference_golden(name, cuda_device):
    g = _golden(name)
    if g is None:
        pytest.skip("golden not generated yet")
    res, st = _ours(name, cuda_device)
    vis = g["radii"] > 0
    m = {}      # every metric is collected first (and written to gpurun_out/ for the parity table in DESIGN.md), then asserted
    m["radii_equal"] = float((res["radii"] == g["radii"]).mean())
    m["tiles_touched_equal"] = float((st["tiles_touched"].astype(np.uint32) == g["tiles_touched"]).mean())
    m["num_rendered"] = [int(st["num_rendered"]), int(g["num_rendered"])]
    same_R = st["num_rendered"] == int(g["num_rendered"])
    m["point_list_equal"] = float((st["point_list"].astype(np.uint32) == g["point_list"]).mean()) if same_R else 0.0
    m["keys_sorted_equal"] = float((st["keys_sorted"].astype(np.uint64) == g["keys_sorted"]).mean()) if same_R else 0.0
    m["ranges_equal"] = float((st["ranges"].astype(np.uint32) == g["ranges"]).mean())
    m["n_contrib_equal"] = float((st["n_contrib"].astype(np.uint32) == g["n_contrib"]).mean())
    m["means2D_maxabs"] = float(np.abs(st["means2D"][vis] - g["means2D_pix"][vis]).max())
    m["conic_rel"] = util.rel_err(st["conic_opacity"][vis], g["conic_opacity"][vis])
    # with precomputed colours the reference leaves its rgb scratch uninitialised (forward.cu:247-253): nothing to compare
    m["rgb_rel"] = util.rel_err(st["rgb"][vis], g["rgb"][vis]) if "g_shs" in res else 0.0
    for k in ("color", "depth", "alpha"):
        m[k + "_rel"] = util.rel_err(res[k], g[k])
    for k in GRAD_KEYS:
        if k in res:
            m[k + "_rel"] = util.rel_err(res[k], g[k])
    os.makedirs(os.path.join(util.ROOT, "gpurun_out"), exist_ok=True)
    import json
    with open(os.path.join(util.ROOT, "gpurun_out", f"gs3d_parity_{name}.json"), "w") as f:
        json.dump(m, f, indent=1)
    print(name, json.dumps(m))
    # tile lists: bit-exact (radii, tiles, depth-sorted instance list, ranges)
    for k in ("radii_equal", "tiles_touched_equal", "point_list_equal", "keys_sorted_equal", "ranges_equal"):
        assert m[k] == 1.0, (k, m)
    assert m["num_rendered"][0] == m["num_rendered"][1], m
    assert m["n_contrib_equal"] > 0.9995, m
    assert m["means2D_maxabs"] < 1e-3 and m["conic_rel"] < 1e-5 and m["rgb_rel"] < 1e-6, m
    for k in ("color", "depth", "alpha"):
        assert m[k + "_rel"] < RTOL_MAPS, (k, m)
    # G1 (camera inside the cloud, screen-filling Gaussians, heavy cancellation): the reference's own fp32 gradients sit 6e-4
    # from the float64 oracle (tests/test_gs3d_cpu.py), ours 8e-4; the two fp32 implementations differ by the same amount
    tol = 1.5e-3 if name == "G1" else RTOL_GRADS
    for k in GRAD_KEYS:
        if k in res:
            assert m[k + "_rel"] < tol, (k, m)


def test_matches_live_reference_at_scale(cuda_device):
    """100 k Gaussians at 800x800 (the C2 size): ours against the unmodified reference extension on the same inputs, plus
    size-independent properties — alpha = sum of weights <= 1, colour of an all-background pixel = background, radii == 0
    exactly for the Gaussians behind the near plane."""
    from d2gs_b200 import synthetic as syn
    import diff_gaussian_rasterization as ours
    rng = np.random.default_rng(77)
    P, W, H = 100_000, 800, 800
    d = rng.normal(size=(P, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    xyz = (d * rng.uniform(size=(P, 1)) ** (1 / 3)).astype(np.float32)
    xyz[:50] += np.array([0, 0, 0], np.float32)
    scales = np.exp(np.log(0.006) + 0.5 * rng.normal(size=(P, 3))).astype(np.float32)
    q = rng.normal(size=(P, 4)); q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    cam = syn.fibonacci_cameras(100, W, H)[37]
    case = dict(inputs=dict(means3D=xyz, opacities=rng.uniform(0.05, 0.95, size=(P, 1)).astype(np.float32),
                            shs=np.concatenate([rng.normal(size=(P, 1, 3)), 0.1 * rng.normal(size=(P, 15, 3))], 1).astype(np.float32),
                            sh_degree=3, scales=scales, rotations=q, scale_modifier=1.0, bg=np.array([0.3, 0.1, 0.7], np.float32),
                            viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                            tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, H=H, W=W),
                g_color=rng.normal(size=(3, H, W)).astype(np.float32), g_depth=rng.normal(size=(1, H, W)).astype(np.float32),
                g_alpha=rng.normal(size=(1, H, W)).astype(np.float32))
    res = run_module(ours, case, cuda_device)
    assert res["alpha"].max() <= 1.0 + 1e-5 and res["alpha"].min() >= 0.0
    empty = res["alpha"][0] == 0
    assert empty.any()
    for ch in range(3):
        assert np.all(res["color"][ch][empty] == case["inputs"]["bg"][ch])
    ref = load_reference_dgr()
    if ref is None:
        pytest.skip("oracle/_ref/diff_gaussian_rasterization not present")
    want = run_module(ref, case, cuda_device)
    assert np.array_equal(res["radii"], want["radii"])
    for k in ("color", "depth", "alpha"):
        assert util.rel_err(res[k], want[k]) < RTOL_MAPS, (k, util.rel_err(res[k], want[k]))
    for k in GRAD_KEYS:
        if k in res:
            assert util.rel_err(res[k], want[k]) < RTOL_GRADS, (k, util.rel_err(res[k], want[k]))


def test_api_errors_and_empty_scene(cuda_device):
    import diff_gaussian_rasterization as ours
    dev = cuda_device
    c = CASES["G2"]["inputs"]
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
    rs = ours.GaussianRasterizationSettings(48, 64, float(c["tanfovx"]), float(c["tanfovy"]), t(c["bg"]), 1.0, t(c["viewmatrix"]),
                                            t(c["projmatrix"]), 1, t(c["campos"]), False, False)
    r = ours.GaussianRasterizer(rs)
    x, o = t(c["means3D"]), t(c["opacities"])
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=x, means2D=torch.zeros_like(x), opacities=o, scales=t(c["scales"]), rotations=t(c["rotations"]))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=torch.zeros_like(x), opacities=o, shs=t(c["shs"]))
    e = torch.zeros((0, 3), device=dev)
    color, radii, depth, alpha = r(means3D=e, means2D=e, opacities=torch.zeros((0, 1), device=dev), shs=torch.zeros((0, 16, 3), device=dev),
                                   scales=e, rotations=torch.zeros((0, 4), device=dev))
    assert color.shape == (3, 48, 64) and float(color.abs().max()) == 0 and radii.numel() == 0 and float(alpha.abs().max()) == 0
    vis = r.markVisible(x)
    assert vis.dtype == torch.bool and vis.shape == (x.shape[0],)


@pytest.mark.parametrize("name", ["G0", "G2"])
def test_binning_modes_are_identical(name, cuda_device):
    """The gs3d path bins like the surfel path (round 2): per-tile buckets + per-tile sort (default) or the reference's global
    radix sort (`tile_sort = 0`), with the instance count read back (default) or deferred (no host synchronisation).  Tile
    lists, ranges, n_contrib and images are bit-identical in all three settings; gradients agree to the order of the sums."""
    import diff_gaussian_rasterization as ours
    from d2gs_b200 import _lib, gs3d, raster
    out = {}
    try:
        for mode in ("global", "tile", "deferred"):
            _lib.set_option("tile_sort", 0 if mode == "global" else 1)
            raster.set_deferred_count(mode == "deferred", warmup=1, margin=1.5)
            if mode == "deferred":
                run_module(ours, CASES[name], cuda_device)              # one synchronous frame establishes the capacity
            res = run_module(ours, CASES[name], cuda_device)
            ctx = gs3d.LAST_CONTEXT
            st = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in gs3d.export_state().items()}
            out[mode] = (res, st, ctx.num_rendered, ctx.R)
    finally:
        _lib.set_option("tile_sort", 1)
        raster.set_deferred_count(False)
    ref, sref, Rref, _ = out["global"]
    assert out["deferred"][3] > out["deferred"][2] == Rref > 0          # the deferred frame ran on capacity, not on the count
    for mode in ("tile", "deferred"):
        res, st, R, _ = out[mode]
        assert R == Rref
        assert np.array_equal(res["radii"], ref["radii"]) and np.array_equal(st["ranges"], sref["ranges"])
        assert np.array_equal(st["point_list"][:R], sref["point_list"][:R]) and np.array_equal(st["keys_sorted"][:R], sref["keys_sorted"][:R])
        assert np.array_equal(st["n_contrib"], sref["n_contrib"])
        for k in ("color", "depth", "alpha"):
            assert np.array_equal(res[k], ref[k]), (mode, k)
        for k in GRAD_KEYS:
            if k in res:
                assert util.rel_err(res[k], ref[k]) < 2e-5, (mode, k)


def test_deferred_overflow_is_loud_and_capturable(cuda_device):
    """(1) A deferred frame whose instance count exceeds its slots renders NaN and the next call raises; (2) with no
    synchronisation left, forward + backward of the gs3d op replays as a CUDA graph."""
    import diff_gaussian_rasterization as ours
    from d2gs_b200 import _lib, gs3d, raster, synthetic as syn
    rng = np.random.default_rng(5)
    P, W, H = 30_000, 400, 304
    d = rng.normal(size=(P, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    cam = syn.fibonacci_cameras(16, W, H)[5]
    q = rng.normal(size=(P, 4))
    case = dict(inputs=dict(means3D=(d * rng.uniform(size=(P, 1)) ** (1 / 3)).astype(np.float32),
                            opacities=rng.uniform(0.05, 0.95, size=(P, 1)).astype(np.float32),
                            shs=np.concatenate([rng.normal(size=(P, 1, 3)), 0.1 * rng.normal(size=(P, 15, 3))], 1).astype(np.float32),
                            sh_degree=3, scales=np.exp(np.log(0.01) + 0.5 * rng.normal(size=(P, 3))).astype(np.float32),
                            rotations=(q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32), scale_modifier=1.0,
                            bg=np.array([0.3, 0.1, 0.7], np.float32), viewmatrix=cam.world_view_transform,
                            projmatrix=cam.full_proj_transform, campos=cam.camera_center, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, H=H, W=W),
                g_color=rng.normal(size=(3, H, W)).astype(np.float32), g_depth=rng.normal(size=(1, H, W)).astype(np.float32),
                g_alpha=rng.normal(size=(1, H, W)).astype(np.float32))       # tens of thousands of instances: far above the 4096-slot floor
    try:
        raster.set_deferred_count(True, warmup=1, margin=1.5)
        run_module(ours, case, cuda_device)
        want = run_module(ours, case, cuda_device)
        assert gs3d.LAST_CONTEXT.num_rendered > 20_000
        key = [k for k in raster._TRACK if k[0] == "gs3d"][-1]
        raster._TRACK[key].poll(key, block=True)                        # fold in the pending counts, then ...
        raster._TRACK[key].max_R = 10                                   # ... pretend the scene used to be tiny
        bad = run_module(ours, case, cuda_device)
        assert np.isnan(bad["color"]).all()
        torch.cuda.synchronize()
        with pytest.raises(_lib.D2gsError, match="deferred-count"):
            run_module(ours, case, cuda_device)
        again = run_module(ours, case, cuda_device)                     # the capacity has been raised
        assert np.array_equal(again["color"], want["color"])
        # CUDA graph: static inputs, capture fwd + bwd once, replay
        dev = cuda_device
        T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=dev)
        inp = case["inputs"]
        xyz = T(inp["means3D"]).requires_grad_(True)
        rs = ours.GaussianRasterizationSettings(image_height=int(inp["H"]), image_width=int(inp["W"]), tanfovx=float(inp["tanfovx"]),
                                                tanfovy=float(inp["tanfovy"]), bg=T(inp["bg"]), scale_modifier=float(inp["scale_modifier"]),
                                                viewmatrix=T(inp["viewmatrix"]), projmatrix=T(inp["projmatrix"]), sh_degree=int(inp["sh_degree"]),
                                                campos=T(inp["campos"]), prefiltered=False, debug=False)
        kw = dict(means2D=torch.zeros_like(xyz), opacities=T(inp["opacities"]),
                  scales=T(inp["scales"]) if inp.get("scales") is not None else None,
                  rotations=T(inp["rotations"]) if inp.get("rotations") is not None else None)
        if inp.get("shs") is not None:
            kw["shs"] = T(inp["shs"])
        else:
            kw["colors_precomp"] = T(inp["colors_precomp"])
        if inp.get("scales") is None:
            kw.pop("scales"); kw.pop("rotations"); kw["cov3D_precomp"] = T(inp["cov3D_precomp"])
        gcol = T(case["g_color"])

        def step():
            color, radii, depth, alpha = ours.GaussianRasterizer(rs)(means3D=xyz, **kw)
            (g,) = torch.autograd.grad((color * gcol).sum() + depth.sum(), xyz)
            return color, g
        s = torch.cuda.Stream(dev)
        with torch.cuda.stream(s):
            c0, g0 = step()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                c1, g1 = step()
            graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(c0, c1) and util.rel_err(g1.cpu().numpy(), g0.cpu().numpy()) < 2e-5
    finally:
        raster.set_deferred_count(False)
