"""Generates tests/golden/knn_golden.npz by running the UNMODIFIED reference simple-knn extension (oracle/_ref/simple_knn,
built by oracle/build_ref_aux.sh from /root/reference/submodules/simple-knn) on a B200:
    gpurun -- python tests/golden/make_knn_golden.py      -> gpurun_out/knn_golden.npz
The file is then copied into tests/golden/ and committed; nothing at test time reads /root/reference.
Cases: the seeded point sets of knn_cases() below (uniform ball, clustered, duplicated points, a flat sheet, tiny P)."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def knn_cases():
    """name -> (P,3) float32.  Shared by the golden generator and the tests (the inputs are NOT stored, only re-drawn)."""
    rng = np.random.default_rng(20261017)
    cases = {}
    d = rng.normal(size=(5000, 3)); r = rng.uniform(size=(5000, 1)) ** (1 / 3)
    cases["ball5000"] = (d / np.linalg.norm(d, axis=1, keepdims=True) * r).astype(np.float32)
    centres = rng.normal(size=(20, 3))
    cases["clusters4099"] = (centres[rng.integers(0, 20, 4099)] + 0.02 * rng.normal(size=(4099, 3))).astype(np.float32)
    base = rng.uniform(-1, 1, size=(1500, 3)).astype(np.float32)
    cases["dups3000"] = np.concatenate([base, base[rng.permutation(1500)]], 0)          # every point has an exact twin
    sheet = rng.uniform(-2, 2, size=(2500, 3)).astype(np.float32); sheet[:, 2] = 0.25
    cases["sheet2500"] = sheet
    cases["tiny7"] = rng.normal(size=(7, 3)).astype(np.float32)
    cases["four"] = rng.normal(size=(4, 3)).astype(np.float32)
    return cases


def load_reference_knn():
    d = os.path.join(ROOT, "oracle", "_ref", "simple_knn")
    so = [f for f in os.listdir(d) if f.startswith("_C") and f.endswith(".so")] if os.path.isdir(d) else []
    if not so:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("ref_simple_knn._C", os.path.join(d, so[0]))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:
        print("reference simple_knn unavailable:", e)
        return None
    return mod


if __name__ == "__main__":
    import torch
    ref = load_reference_knn()
    assert ref is not None, "oracle/_ref/simple_knn missing: run oracle/build_ref_aux.sh in the build container first"
    out = {}
    for name, pts in knn_cases().items():
        out[name] = ref.distCUDA2(torch.as_tensor(pts, device="cuda")).cpu().numpy()
        print(name, pts.shape, float(out[name].mean()))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "knn_golden.npz"), **out)
