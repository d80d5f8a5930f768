"""Generates tests/golden/epilogue_golden.npz with the REFERENCE's own ``depth_to_normal`` / ``depths_to_points``
(/root/reference/utils/point_utils.py:9-38, imported in the build container; ``Tensor.cuda()`` made a no-op and the unused
``cv2`` / ``matplotlib`` imports stubbed so that it runs on CPU tensors).  These two functions are the part of render()'s
image-space epilogue (gaussian_renderer/__init__.py:172-207) that lives outside render() itself; the restatement in
oracle/reference_pipeline.py — the yardstick of the fused epilogue kernel — is checked against this file.
Committed together with its output; nothing at test time reads /root/reference."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "dynamic-2dgs_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def epilogue_case(name):
    """Seeded camera + depth map (shared with the tests)."""
    from d2gs_b200 import synthetic as syn
    W, H, idx, seed = {"a": (96, 64, 3, 5), "b": (50, 70, 6, 6)}[name]
    cam = syn.fibonacci_cameras(8, W, H)[idx]
    g = torch.Generator().manual_seed(seed)
    depth = 3.0 + torch.rand(1, H, W, generator=g) + 0.3 * torch.sin(torch.linspace(0, 9, W))[None, None, :]
    if name == "b":
        depth[:, 10:20, 5:15] = 0.0                       # pixels nothing was rendered to
    view = types.SimpleNamespace(world_view_transform=torch.as_tensor(cam.world_view_transform), image_width=W, image_height=H,
                                 FoVx=cam.FoVx, FoVy=cam.FoVy)
    return view, depth


if __name__ == "__main__":
    for m in ("cv2", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, "/root/reference")
    from utils.point_utils import depth_to_normal, depths_to_points      # noqa: E402  (reference code, read-only)
    out = {}
    for name in ("a", "b"):
        view, depth = epilogue_case(name)
        normal, points = depth_to_normal(view, depth)
        out[f"{name}_normal"] = normal.numpy()
        out[f"{name}_points"] = depths_to_points(view, depth).numpy()
        print(name, normal.shape, float(normal.abs().mean()))
    np.savez_compressed(os.path.join(HERE, "epilogue_golden.npz"), **out)
