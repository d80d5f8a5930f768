"""Shared helpers of the test-suite: seeded inputs in the reference's calling convention."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dynamic-2dgs_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

from d2gs_b200 import synthetic as syn  # noqa: E402


def raster_inputs(cfg_name="T0", cam_index=3, n_cams=8, bg=(0.1, 0.2, 0.3), s_med=None, seed=None, sh_degree=3):
    """numpy dict: activated surfel parameters + camera + settings for one seeded view."""
    cfg = dict(syn.CONFIGS[cfg_name])
    if s_med is not None:
        cfg["s_med"] = s_med
    if seed is not None:
        cfg["seed"] = seed
    sc = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"], sh_degree=3)
    cam = syn.fibonacci_cameras(n_cams, cfg["W"], cfg["H"])[cam_index]
    act = syn.activated(sc)
    kw = dict(bg=np.asarray(bg, np.float32), viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
              campos=cam.camera_center, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, image_height=cam.image_height,
              image_width=cam.image_width, sh_degree=sh_degree)
    return act, kw


def upstream_grads(H, W, seed=0):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(3, H, W)).astype(np.float32), rng.normal(size=(8, H, W)).astype(np.float32)


def load_reference_ext():
    """The UNMODIFIED reference op built by oracle/build_ref.sh into oracle/_ref (None if absent)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "diff_surfel_rasterization")):
        return None
    spec = importlib.util.spec_from_file_location(
        "ref_diff_surfel_rasterization", os.path.join(ref_dir, "diff_surfel_rasterization", "__init__.py"),
        submodule_search_locations=[os.path.join(ref_dir, "diff_surfel_rasterization")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_diff_surfel_rasterization"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception as e:  # no GPU / ABI mismatch
        print("reference ext unavailable:", e)
        return None
    return mod


def settings_for(mod, kw, device, debug=False):
    import torch
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    return mod.GaussianRasterizationSettings(
        image_height=int(kw["image_height"]), image_width=int(kw["image_width"]), tanfovx=float(kw["tanfovx"]),
        tanfovy=float(kw["tanfovy"]), bg=t(kw["bg"]), scale_modifier=1.0, viewmatrix=t(kw["viewmatrix"]),
        projmatrix=t(kw["projmatrix"]), sh_degree=int(kw["sh_degree"]), campos=t(kw["campos"]), prefiltered=False,
        debug=debug)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rel_err_trimmed(a, b, trim=1e-3):
    """Norm-wise relative error after dropping the `trim` fraction of elements with the largest |a-b|.
    For maps that are DISCONTINUOUS in the rasterizer inputs (median depth switches contributor when the transmittance
    crosses 0.5; a radius flips by a pixel): a handful of flipped pixels would otherwise dominate the L2 norm."""
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    d = np.abs(a - b)
    k = int(np.ceil(trim * d.size))
    if 0 < k < d.size:
        d = np.partition(d, d.size - k)[: d.size - k]
    return float(np.linalg.norm(d) / max(np.linalg.norm(b), 1e-30))
