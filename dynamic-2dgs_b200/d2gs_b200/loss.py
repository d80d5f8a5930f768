"""Fused photometric loss of the training step (libd2gs.so: d2gs_loss_forward/backward).

Mirrors the reference's loss code so a trainer can swap it in:

    Ll1 = l1_loss(image, gt); loss = (1 - l) * Ll1 + l * (1 - ssim(image, gt)) + normal_loss + dist_loss
        (utils/loss_utils.py:18-19,33-76, train_gui.py:292-313)

becomes ``loss = surfel_loss(image, gt, rend_normal, surf_normal, rend_dist, lambda_dssim, lambda_normal, lambda_dist)``:
one kernel forward (separable 11x11 Gaussian windows in shared memory, all four terms and their sums) and one backward,
instead of 5 grouped conv2d + ~25 elementwise/reduction kernels each way.  ``l1_loss`` / ``ssim`` with the reference's
names are provided on top of the same kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _img(t: torch.Tensor, c: int, name: str) -> torch.Tensor:
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3 or t.shape[0] != c:
        raise ValueError(f"{name} must have shape ({c},H,W), got {tuple(t.shape)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the fused loss has no CPU path)")
    return t.detach().float().contiguous()


class _SurfelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, gt, rend_normal, surf_normal, rend_dist, l_dssim, l_normal, l_dist):
        L = _lib.lib()
        img, tgt = _img(image, 3, "image"), _img(gt, 3, "gt")
        if img.shape != tgt.shape:
            raise ValueError("image and gt differ in shape")
        dev = img.device
        H, W = int(img.shape[1]), int(img.shape[2])
        use_n = rend_normal is not None and surf_normal is not None and l_normal != 0.0
        use_d = rend_dist is not None and l_dist != 0.0
        rn = _img(rend_normal, 3, "rend_normal") if use_n else None
        sn = _img(surf_normal, 3, "surf_normal") if use_n else None
        rd = _img(rend_dist, 1, "rend_dist") if use_d else None
        need_grad = any(t is not None and torch.is_tensor(t) and t.requires_grad for t in (image, rend_normal, surf_normal, rend_dist))
        nbytes = C.c_size_t()
        _lib.check(L.d2gs_loss_workspace(W, H, C.byref(nbytes)), "d2gs_loss_workspace")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=dev)
        out = torch.empty((5,), dtype=torch.float32, device=dev)
        a = _lib.LossArgs()
        a.width, a.height = W, H
        a.image, a.gt = img.data_ptr(), tgt.data_ptr()
        a.rend_normal = rn.data_ptr() if use_n else None
        a.surf_normal = sn.data_ptr() if use_n else None
        a.rend_dist = rd.data_ptr() if use_d else None
        a.lambda_dssim, a.lambda_normal, a.lambda_dist = float(l_dssim), float(l_normal), float(l_dist)
        a.out, a.workspace, a.workspace_bytes = out.data_ptr(), ws.data_ptr(), nbytes.value
        a.save_for_backward = int(need_grad)
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_loss_forward(C.byref(a), _stream(dev)), "d2gs_loss_forward")
        ctx.cfg = (W, H, float(l_dssim), float(l_normal), float(l_dist), use_n, use_d,
                   tuple(image.shape), None if rend_normal is None else tuple(rend_normal.shape),
                   None if surf_normal is None else tuple(surf_normal.shape), None if rend_dist is None else tuple(rend_dist.shape))
        ctx.save_for_backward(img, tgt, rn, sn, rd, ws)
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        L = _lib.lib()
        W, H, l_dssim, l_normal, l_dist, use_n, use_d, s_img, s_rn, s_sn, s_rd = ctx.cfg
        img, tgt, rn, sn, rd, ws = ctx.saved_tensors
        dev = img.device
        up = g_loss.detach().float().reshape(1).contiguous()
        g_img = torch.empty_like(img)
        want_rn, want_sn, want_rd = (use_n and ctx.needs_input_grad[2]), (use_n and ctx.needs_input_grad[3]), (use_d and ctx.needs_input_grad[4])
        g_rn = torch.empty_like(rn) if want_rn else None
        g_sn = torch.empty_like(sn) if want_sn else None
        g_rd = torch.empty_like(rd) if want_rd else None
        a = _lib.LossArgs()
        a.width, a.height = W, H
        a.image, a.gt = img.data_ptr(), tgt.data_ptr()
        a.rend_normal = rn.data_ptr() if use_n else None
        a.surf_normal = sn.data_ptr() if use_n else None
        a.rend_dist = rd.data_ptr() if use_d else None
        a.lambda_dssim, a.lambda_normal, a.lambda_dist = l_dssim, l_normal, l_dist
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.upstream = up.data_ptr()
        a.g_image = g_img.data_ptr()
        a.g_rend_normal = g_rn.data_ptr() if want_rn else None
        a.g_surf_normal = g_sn.data_ptr() if want_sn else None
        a.g_rend_dist = g_rd.data_ptr() if want_rd else None
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_loss_backward(C.byref(a), _stream(dev)), "d2gs_loss_backward")
        rs = lambda g, shp: None if g is None else g.reshape(shp)
        return (rs(g_img, s_img) if ctx.needs_input_grad[0] else None, None, rs(g_rn, s_rn), rs(g_sn, s_sn), rs(g_rd, s_rd),
                None, None, None)


def surfel_loss(image, gt, rend_normal=None, surf_normal=None, rend_dist=None, lambda_dssim: float = 0.2,
                lambda_normal: float = 0.0, lambda_dist: float = 0.0, return_parts: bool = False):
    """Total training loss (scalar tensor).  With ``return_parts`` also the detached (5,) tensor
    [loss, L1, SSIM, weighted normal term, weighted distortion term]."""
    loss, parts = _SurfelLoss.apply(image, gt, rend_normal, surf_normal, rend_dist, lambda_dssim, lambda_normal, lambda_dist)
    return (loss, parts) if return_parts else loss


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19 on the fused kernel (lambda_dssim = 0 leaves exactly the L1 term)."""
    return surfel_loss(network_output, gt, lambda_dssim=0.0)


def ssim(img1, img2, window_size: int = 11, size_average: bool = True):
    """utils/loss_utils.py:45-76 (window 11, mean over all elements) on the fused kernel: 1 - loss at lambda_dssim = 1."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("the fused SSIM implements the reference's training configuration (window 11, mean)")
    return 1.0 - surfel_loss(img1, img2, lambda_dssim=1.0)
