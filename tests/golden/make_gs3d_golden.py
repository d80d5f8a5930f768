"""Generates tests/golden/gs3d_golden_<case>.npz by running the UNMODIFIED reference depth/alpha rasterizer
(oracle/_ref/diff_gaussian_rasterization, built by oracle/build_ref_aux.sh from
/root/reference/submodules/diff-gaussian-rasterization) on a B200:
    gpurun -- python tests/golden/make_gs3d_golden.py      -> gpurun_out/gs3d_golden_<case>.npz
Outputs, every intermediate of its geometry / binning / image buffers (decoded with the bump-allocation rule of
DGR/cuda_rasterizer/rasterizer_impl.h:21-28 + rasterizer_impl.cu:154-194) and all eight gradients are recorded; inputs are
re-drawn from gs3d_cases() by the tests, not stored.  Nothing at test time reads /root/reference."""
import importlib.util
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "dynamic-2dgs_b200"), ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def gs3d_cases():
    """name -> dict(inputs for the rasterizer as numpy arrays, upstream gradients).  Seeded; shared by generator and tests."""
    from d2gs_b200 import synthetic as syn
    cases = {}

    def scene(P, seed, s_med, anis):
        rng = np.random.default_rng(seed)
        d = rng.normal(size=(P, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        xyz = (d * rng.uniform(size=(P, 1)) ** (1 / 3)).astype(np.float32)
        scales = np.exp(math.log(s_med) + anis * rng.normal(size=(P, 3))).astype(np.float32)
        q = rng.normal(size=(P, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        opac = rng.uniform(0.02, 0.98, size=(P, 1)).astype(np.float32)
        shs = np.concatenate([rng.normal(size=(P, 1, 3)), 0.15 * rng.normal(size=(P, 15, 3))], 1).astype(np.float32)
        return rng, xyz, scales, q.astype(np.float32), opac, shs

    def grads(rng, H, W):
        return dict(g_color=rng.normal(size=(3, H, W)).astype(np.float32), g_depth=rng.normal(size=(1, H, W)).astype(np.float32),
                    g_alpha=rng.normal(size=(1, H, W)).astype(np.float32))

    def cam_kw(cam):
        return dict(viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                    tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, H=cam.image_height, W=cam.image_width)

    # G0: SH degree 3, scale/rotation, ragged image (100x72: partial tiles on both axes), non-zero background
    rng, xyz, scales, q, opac, shs = scene(1500, 101, 0.035, 0.6)
    cam = syn.fibonacci_cameras(9, 100, 72)[4]
    cases["G0"] = dict(inputs=dict(means3D=xyz, opacities=opac, shs=shs, sh_degree=3, scales=scales, rotations=q, scale_modifier=1.0,
                                   bg=np.array([0.2, 0.5, 0.1], np.float32), **cam_kw(cam)), **grads(rng, 72, 100))
    # G1: precomputed colours + precomputed covariance, camera INSIDE the cloud (near-plane culls, means beyond the
    #     1.3 tan(fov) clamp), opaque Gaussians (early termination), degree-independent path, scale_modifier ignored
    rng, xyz, scales, q, opac, shs = scene(900, 202, 0.06, 0.4)
    opac = np.clip(opac * 1.6, 0.0, 0.999).astype(np.float32)
    cam = syn.look_at_camera((0.35, -0.2, 0.1), 80, 64, target=(-0.4, 0.5, 0.0))
    Rm = np.zeros((900, 3, 3)); r, x, y, z = [q[:, i].astype(np.float64) for i in range(4)]
    Rm[:, 0] = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1)
    Rm[:, 1] = np.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1)
    Rm[:, 2] = np.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)
    Sg = Rm @ (scales.astype(np.float64)[:, :, None] ** 2 * Rm.transpose(0, 2, 1))
    cov6 = np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], -1).astype(np.float32)
    cases["G1"] = dict(inputs=dict(means3D=xyz, opacities=opac, colors_precomp=rng.uniform(0, 1, size=(900, 3)).astype(np.float32),
                                   cov3D_precomp=cov6, scale_modifier=1.0, bg=np.zeros(3, np.float32), sh_degree=0, **cam_kw(cam)),
                       **grads(rng, 64, 80))
    # G2: SH degree 1 of 16 stored coefficients, scale_modifier 0.7, quaternions NOT unit (the variant uses them as given)
    rng, xyz, scales, q, opac, shs = scene(700, 303, 0.05, 0.5)
    q = (q * rng.uniform(0.8, 1.2, size=(700, 1))).astype(np.float32)
    cam = syn.fibonacci_cameras(5, 64, 48)[1]
    cases["G2"] = dict(inputs=dict(means3D=xyz, opacities=opac, shs=shs, sh_degree=1, scales=scales, rotations=q, scale_modifier=0.7,
                                   bg=np.ones(3, np.float32), **cam_kw(cam)), **grads(rng, 48, 64))
    return cases


def load_reference_dgr():
    d = os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization")
    if not os.path.isfile(os.path.join(d, "__init__.py")):
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("ref_diff_gaussian_rasterization", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_diff_gaussian_rasterization"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as e:
        print("reference diff_gaussian_rasterization unavailable:", e)
        return None
    return mod


def run_module(mod, case, device, keep_ctx=False):
    """Runs fwd+bwd of a rasterizer package (the reference's or the drop-in) on one case; returns numpy outputs + gradients."""
    import torch
    inp = case["inputs"]
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    settings = mod.GaussianRasterizationSettings(
        image_height=int(inp["H"]), image_width=int(inp["W"]), tanfovx=float(inp["tanfovx"]), tanfovy=float(inp["tanfovy"]),
        bg=t(inp["bg"]), scale_modifier=float(inp["scale_modifier"]), viewmatrix=t(inp["viewmatrix"]), projmatrix=t(inp["projmatrix"]),
        sh_degree=int(inp["sh_degree"]), campos=t(inp["campos"]), prefiltered=False, debug=False)
    leaves = {k: t(inp[k]).requires_grad_(True) for k in ("means3D", "opacities", "shs", "colors_precomp", "scales", "rotations", "cov3D_precomp")
              if inp.get(k) is not None}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    color, radii, depth, alpha = mod.GaussianRasterizer(settings)(
        means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves.get("shs"),
        colors_precomp=leaves.get("colors_precomp"), scales=leaves.get("scales"), rotations=leaves.get("rotations"),
        cov3D_precomp=leaves.get("cov3D_precomp"))
    loss = (color * t(case["g_color"])).sum() + (depth * t(case["g_depth"])).sum() + (alpha * t(case["g_alpha"])).sum()
    # the autograd node of the op is its ctx: saved tensors must be read before backward() frees them
    kept = (tuple(color.grad_fn.saved_tensors), int(getattr(color.grad_fn, "num_rendered", 0))) if keep_ctx else None
    loss.backward()
    torch.cuda.synchronize()
    out = dict(color=color, radii=radii, depth=depth, alpha=alpha, g_means2D=means2D.grad)
    for k, v in leaves.items():
        out["g_" + k] = v.grad
    res = {k: v.detach().cpu().numpy() for k, v in out.items() if v is not None}
    return (res, kept) if keep_ctx else res


if __name__ == "__main__":
    import torch
    from make_golden import _carve
    ref = load_reference_dgr()
    assert ref is not None, "oracle/_ref/diff_gaussian_rasterization missing: run oracle/build_ref_aux.sh first"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name, case in gs3d_cases().items():
        res, (saved, R) = run_module(ref, case, "cuda", keep_ctx=True)
        P, H, W = case["inputs"]["means3D"].shape[0], int(case["inputs"]["H"]), int(case["inputs"]["W"])
        # saved: colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img, alpha
        geom, binning, img = saved[7], saved[8], saved[9]
        g = _carve(geom, [("depths", np.float32, P, 1), ("clamped", np.uint8, P, 3), ("internal_radii", np.int32, P, 1),
                          ("means2D", np.float32, P, 2), ("cov3D", np.float32, P, 6), ("conic_opacity", np.float32, P, 4),
                          ("rgb", np.float32, P, 3), ("tiles_touched", np.uint32, P, 1)])
        im = _carve(img, [("n_contrib", np.uint32, H * W, 1), ("ranges", np.uint32, H * W, 2)])
        b = _carve(binning, [("point_list", np.uint32, R, 1), ("point_list_unsorted", np.uint32, R, 1), ("keys_sorted", np.uint64, R, 1)])
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        res.update(num_rendered=np.int64(R), means2D_pix=g["means2D"], depths=g["depths"], cov3D=g["cov3D"],
                   conic_opacity=g["conic_opacity"], rgb=g["rgb"], clamped=g["clamped"], tiles_touched=g["tiles_touched"],
                   n_contrib=im["n_contrib"].reshape(H, W), ranges=im["ranges"][:tiles], point_list=b["point_list"],
                   keys_sorted=b["keys_sorted"])
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"gs3d_golden_{name}.npz"), **res)
        print(name, "P", P, "R", R, "visible", int((res["radii"] > 0).sum()), "mean alpha", float(res["alpha"].mean()))
