"""Generates tests/golden/camera_golden.npz: the camera matrices the REFERENCE builds (scene/cameras.py:49-59 with
utils/graphics_utils.py:42-77 getWorld2View2 / getProjectionMatrix, imported from /root/reference in the build container)
for the extrinsics and fields of view of three seeded synthetic cameras.  d2gs_b200/synthetic.py must reproduce them:
its cameras feed every parity test and the bench.  Committed together with its output."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "dynamic-2dgs_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def camera_cases():
    from d2gs_b200 import synthetic as syn
    return {"fib3": syn.fibonacci_cameras(8, 128, 96)[3], "fib17": syn.fibonacci_cameras(100, 800, 800)[17],
            "inside": syn.look_at_camera((0.35, -0.2, 0.1), 80, 64, target=(-0.4, 0.5, 0.0))}


if __name__ == "__main__":
    sys.path.insert(0, "/root/reference")
    from utils.graphics_utils import getProjectionMatrix, getWorld2View2      # noqa: E402  (reference code, read-only)
    out = {}
    for name, cam in camera_cases().items():
        w2c = cam.world_view_transform.T.astype(np.float64)
        R, T = w2c[:3, :3].T, w2c[:3, 3]                                    # the reference stores R transposed (getWorld2View2)
        wv = torch.tensor(getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
        proj = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=cam.FoVx, fovY=cam.FoVy).transpose(0, 1)
        full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
        out[f"{name}_world_view_transform"] = wv.numpy()
        out[f"{name}_projection_matrix"] = proj.numpy()
        out[f"{name}_full_proj_transform"] = full.numpy()
        out[f"{name}_camera_center"] = wv.inverse()[3, :3].numpy()
    np.savez_compressed(os.path.join(HERE, "camera_golden.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("fib3")})
