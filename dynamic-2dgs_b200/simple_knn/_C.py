"""``simple_knn._C`` — same name and signature as the reference's pybind module (submodules/simple-knn/ext.cpp:15-17).

distCUDA2(points) -> (P,) float32: mean squared distance of every point to its 3 nearest neighbours, the statistic
``GaussianModel.create_from_pcd`` clamps and takes the log-sqrt of to initialise the surfel scales
(scene/gaussian_model.py:162-163).  Calls d2gs_knn_mean_dist2 of libd2gs.so through ctypes; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from d2gs_b200 import _lib

__all__ = ["distCUDA2"]


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    """submodules/simple-knn/spatial.cu:15-26: (P,3) float32 CUDA points -> (P,) float32 (on the points' device)."""
    if not torch.is_tensor(points) or not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (no CPU path)")
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError(f"distCUDA2: points must have shape (P, 3), got {tuple(points.shape)}")
    L = _lib.lib()
    pts = points.detach().float().contiguous()
    P = int(pts.shape[0])
    means = torch.zeros((P,), dtype=torch.float32, device=pts.device)     # spatial.cu:21: torch::full({P}, 0.0)
    if P == 0:
        return means
    nbytes = C.c_size_t()
    _lib.check(L.d2gs_knn_mean_dist2_workspace(P, C.byref(nbytes)), "d2gs_knn_mean_dist2_workspace")
    ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        _lib.check(L.d2gs_knn_mean_dist2(P, pts.data_ptr(), means.data_ptr(), ws.data_ptr(), nbytes.value, stream),
                   "d2gs_knn_mean_dist2")
    return means
