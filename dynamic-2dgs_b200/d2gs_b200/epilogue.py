"""Fused image-space epilogue of render() (libd2gs.so: d2gs_epilogue_forward/backward).

Replaces gaussian_renderer/__init__.py:172-207 + utils/point_utils.py:9-38 of the reference: alpha, world-space
normals, distortion, nan_to_num'd median depth, unprojected surface points and the alpha-weighted normal of the
depth map — one kernel forward, one backward, no meshgrid / matrix inverse on the host."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class _RenderEpilogue(torch.autograd.Function):
    @staticmethod
    def forward(ctx, allmap, viewmatrix, focal_x, focal_y):
        L = _lib.lib()
        if not allmap.is_cuda:
            raise RuntimeError("allmap must be a CUDA tensor")
        allmap_ = allmap.detach().float().contiguous()
        view_ = viewmatrix.detach().float().contiguous()
        _, H, W = allmap_.shape
        dev = allmap_.device
        mk = lambda c: torch.empty((c, H, W), dtype=torch.float32, device=dev)
        alpha, rend_normal, rend_dist, depth, surf_normal, surf_point = mk(1), mk(3), mk(1), mk(1), mk(3), mk(3)
        a = _lib.EpilogueArgs()
        a.width, a.height, a.allmap, a.viewmatrix = W, H, allmap_.data_ptr(), view_.data_ptr()
        a.focal_x, a.focal_y = float(focal_x), float(focal_y)
        a.alpha, a.rend_normal, a.rend_dist, a.depth = alpha.data_ptr(), rend_normal.data_ptr(), rend_dist.data_ptr(), depth.data_ptr()
        a.surf_normal, a.surf_point = surf_normal.data_ptr(), surf_point.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_epilogue_forward(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "d2gs_epilogue_forward")
        ctx.save_for_backward(allmap_, view_)
        ctx.f = (float(focal_x), float(focal_y))
        return alpha, rend_normal, rend_dist, depth, surf_normal, surf_point

    @staticmethod
    def backward(ctx, g_alpha, g_rn, g_dist, g_depth, g_sn, g_sp):
        L = _lib.lib()
        allmap_, view_ = ctx.saved_tensors
        _, H, W = allmap_.shape
        dev = allmap_.device
        c = lambda g: None if g is None else g.float().contiguous()
        gs = [c(g) for g in (g_alpha, g_rn, g_dist, g_depth, g_sn, g_sp)]
        dA = torch.empty_like(allmap_)
        a = _lib.EpilogueArgs()
        a.width, a.height, a.allmap, a.viewmatrix = W, H, allmap_.data_ptr(), view_.data_ptr()
        a.focal_x, a.focal_y = ctx.f
        p = lambda g: None if g is None else g.data_ptr()
        a.g_alpha, a.g_rend_normal, a.g_rend_dist, a.g_depth, a.g_surf_normal, a.g_surf_point = [p(g) for g in gs]
        a.dL_dallmap = dA.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_epilogue_backward(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "d2gs_epilogue_backward")
        return dA, None, None, None


def render_epilogue(allmap: torch.Tensor, view):
    """(alpha, rend_normal, rend_dist, depth, surf_normal, surf_point) from the rasterizer's 8 planes."""
    W, H = int(view.image_width), int(view.image_height)
    fx = W / (2 * math.tan(view.FoVx / 2.))
    fy = H / (2 * math.tan(view.FoVy / 2.))
    return _RenderEpilogue.apply(allmap, view.world_view_transform, fx, fy)
