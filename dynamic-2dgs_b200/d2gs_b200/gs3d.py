"""Autograd front-end of the 3-D Gaussian rasterizer with depth and alpha outputs (libd2gs.so: d2gs_gs3d_*).

Mirrors ``_RasterizeGaussians`` of the reference's second rasterizer (DGR = submodules/diff-gaussian-rasterization,
DGR/diff_gaussian_rasterization/__init__.py:43-157): same argument order, outputs ``(color (3,H,W), radii (P) int32,
depth (1,H,W), alpha (1,H,W))``, the same eight gradient slots, the same debug-snapshot behaviour.  SURVEY.md §8(f) rank 4.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from . import raster as _raster
from .raster import _opt, _prep, _ptr, _stream_ptr, cpu_deep_copy_tuple

LAST_CONTEXT = None     # parity tests read the intermediates of the most recent forward through export_state()


class Gs3dContext:
    """``R`` = instance slots the binning workspace was laid out for (the C ABI's num_rendered in backward / export_state);
    ``num_rendered`` = the true count (in deferred-count mode reading it waits for the asynchronous copy)."""
    __slots__ = ("P", "D", "M", "W", "H", "R", "geom", "img", "binning", "bg", "view", "proj", "campos", "tanfovx", "tanfovy",
                 "scale_modifier", "debug", "_count", "_count_event", "_count_host")

    @property
    def num_rendered(self) -> int:
        if self._count is None:
            if self._count_event is None:
                torch.cuda.synchronize(self.geom.device)
            else:
                self._count_event.synchronize()
            self._count = int(self._count_host[0]) & 0xffffffff
        return self._count


def _forward(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
    L = _lib.lib()
    means3D = _prep(means3D, "means3D")
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = means3D.device
    sh, colors_precomp = _prep(_opt(sh), "sh"), _prep(_opt(colors_precomp), "colors_precomp")
    opacities = _prep(opacities, "opacities")
    scales, rotations = _prep(_opt(scales), "scales"), _prep(_opt(rotations), "rotations")
    cov3Ds_precomp = _prep(_opt(cov3Ds_precomp), "cov3Ds_precomp")
    P, H, W = int(means3D.shape[0]), int(rs.image_height), int(rs.image_width)
    M = int(sh.shape[1]) if sh is not None else 0
    f32 = dict(dtype=torch.float32, device=dev)
    color = torch.empty((3, H, W), **f32)
    depth = torch.empty((1, H, W), **f32)
    alpha = torch.empty((1, H, W), **f32)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    ctx = Gs3dContext()
    ctx.P, ctx.D, ctx.M, ctx.W, ctx.H, ctx.R = P, int(rs.sh_degree), M, W, H, 0
    ctx.bg, ctx.view = _prep(rs.bg, "bg"), _prep(rs.viewmatrix, "viewmatrix")
    ctx.proj, ctx.campos = _prep(rs.projmatrix, "projmatrix"), _prep(rs.campos, "campos")
    ctx.tanfovx, ctx.tanfovy, ctx.scale_modifier, ctx.debug = float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier), bool(rs.debug)
    g, i, b = C.c_size_t(), C.c_size_t(), C.c_size_t()
    key = ("gs3d", dev.index, P, W, H)
    # Deferred-count mode: the same opt-in switch, warm-up and capacity rule as the surfel rasterizer
    # (raster.set_deferred_count; always used under CUDA-graph capture) — see raster.raster_forward.
    track = _raster._lru_get(_raster._TRACK, key, _raster._CountTrack, _raster._TRACK_MAX, on_evict=lambda k: _raster._R_HINT.pop(k, None))
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        if ctx.debug or P == 0 or track.max_R <= 0:
            raise _lib.D2gsError("CUDA-graph capture of the gs3d rasterizer needs debug=False and at least one eager frame of this "
                                 "(P, width, height) first: the binning capacity comes from observed instance counts")
        deferred = True
    else:
        track.poll(key)
        track.raise_if_overflowed()
        if not track.spare:
            track.spare = [torch.zeros((2,), dtype=torch.int32).pin_memory() for _ in range(4)]
        dcfg = _raster._DEFERRED
        deferred = bool(dcfg["on"]) and P > 0 and not ctx.debug and track.frames >= dcfg["warmup"] and track.max_R > 0
    if deferred:
        hint = int(track.max_R * _raster._DEFERRED["margin"]) + 4096
    else:
        hint = int(_raster._R_HINT.get(key, 4 * P + 1024) * 1.25) + 1024
    _lib.check(L.d2gs_gs3d_workspace(P, W, H, hint, C.byref(g), C.byref(i), C.byref(b)), "d2gs_gs3d_workspace")
    u8 = dict(dtype=torch.uint8, device=dev)
    ctx.geom, ctx.img = torch.empty((g.value,), **u8), torch.empty((i.value,), **u8)
    ctx.binning = torch.empty((b.value,), **u8)
    a = _lib.Gs3dFwdArgs()
    a.P, a.D, a.M, a.width, a.height = P, ctx.D, M, W, H
    a.background, a.means3D, a.shs, a.colors_precomp = _ptr(ctx.bg), _ptr(means3D), _ptr(sh), _ptr(colors_precomp)
    a.opacities, a.scales, a.rotations, a.cov3D_precomp = _ptr(opacities), _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp)
    a.scale_modifier = ctx.scale_modifier
    a.viewmatrix, a.projmatrix, a.campos = _ptr(ctx.view), _ptr(ctx.proj), _ptr(ctx.campos)
    a.tan_fovx, a.tan_fovy = ctx.tanfovx, ctx.tanfovy
    a.prefiltered, a.debug = int(bool(rs.prefiltered)), int(ctx.debug)
    a.out_color, a.out_depth, a.out_alpha, a.radii = _ptr(color), _ptr(depth), _ptr(alpha), _ptr(radii)
    a.geom_buffer, a.geom_bytes = ctx.geom.data_ptr(), ctx.geom.numel()
    a.img_buffer, a.img_bytes = ctx.img.data_ptr(), ctx.img.numel()
    a.binning_buffer, a.binning_bytes = ctx.binning.data_ptr(), ctx.binning.numel()
    R, need = C.c_int64(0), C.c_size_t(0)
    a.num_rendered, a.binning_required = C.pointer(R), C.pointer(need)
    ctx._count, ctx._count_event, ctx._count_host = None, None, None
    if deferred:
        if capturing:
            if not track.spare:
                raise _lib.D2gsError("more than 4 gs3d frames of one (P, width, height) captured without an eager frame in between")
            host = track.spare.pop()
        else:
            host = torch.empty((2,), dtype=torch.int32, pin_memory=True)
        a.binning_capacity, a.num_rendered_async = hint, host.data_ptr()
        ctx._count_host = host
    with torch.cuda.device(dev):
        rc = L.d2gs_gs3d_forward(C.byref(a), _stream_ptr(dev))
        if rc == _lib.D2GS_NEED_BINNING:
            ctx.binning = torch.empty((need.value,), **u8)
            a.binning_buffer, a.binning_bytes, a.resume = ctx.binning.data_ptr(), ctx.binning.numel(), 1
            rc = L.d2gs_gs3d_forward(C.byref(a), _stream_ptr(dev))
        _lib.check(rc, "d2gs_gs3d_forward")
        if deferred and not capturing:
            ctx._count_event = torch.cuda.Event()
            ctx._count_event.record(torch.cuda.current_stream(dev))
    ctx.R = int(R.value)
    if capturing:
        track.captured.append((host, hint))
    elif deferred:
        track.pending.append((ctx._count_event, host, hint))
    else:
        ctx._count = ctx.R
        track.observe(ctx.R)
        _raster._R_HINT[key] = ctx.R
    return ctx, color, depth, alpha, radii, (means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        global LAST_CONTEXT
        args = (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)     # copy before anything can corrupt them (DGR/.../__init__.py:83-90)
            try:
                rctx, color, depth, alpha, radii, kept = _forward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            rctx, color, depth, alpha, radii, kept = _forward(*args)
        ctx.rctx = rctx
        ctx.has = [t is not None for t in kept]
        ctx.save_for_backward(radii, alpha, *[t for t in kept if t is not None])
        ctx.mark_non_differentiable(radii)
        LAST_CONTEXT = rctx
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        L = _lib.lib()
        r = ctx.rctx
        radii, alpha, *rest = ctx.saved_tensors
        it = iter(rest)
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp = (next(it) if h else None for h in ctx.has)
        dev = means3D.device
        P, M, H, W = r.P, r.M, r.H, r.W
        f32 = dict(dtype=torch.float32, device=dev)
        z = lambda t, shape: torch.zeros(shape, **f32) if t is None else t.float().contiguous()
        grad_color, grad_depth, grad_alpha = z(grad_color, (3, H, W)), z(grad_depth, (1, H, W)), z(grad_alpha, (1, H, W))
        g_means2D, g_colors = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32)
        g_opacity, g_means3D = torch.empty((P, 1), **f32), torch.empty((P, 3), **f32)
        g_cov3D, g_sh = torch.empty((P, 6), **f32), torch.empty((P, M, 3), **f32)
        g_scales, g_rot = torch.empty((P, 3), **f32), torch.empty((P, 4), **f32)
        if P:
            scratch = torch.empty((P, 12), **f32)
            a = _lib.Gs3dBwdArgs()
            a.P, a.D, a.M, a.width, a.height, a.num_rendered = P, r.D, M, W, H, r.R
            a.background, a.means3D, a.shs, a.colors_precomp = _ptr(r.bg), _ptr(means3D), _ptr(sh), _ptr(colors_precomp)
            a.scales, a.rotations, a.cov3D_precomp, a.scale_modifier = _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp), r.scale_modifier
            a.viewmatrix, a.projmatrix, a.campos = _ptr(r.view), _ptr(r.proj), _ptr(r.campos)
            a.tan_fovx, a.tan_fovy = r.tanfovx, r.tanfovy
            a.radii, a.out_alpha = _ptr(radii), _ptr(alpha)
            a.geom_buffer, a.binning_buffer, a.img_buffer = r.geom.data_ptr(), r.binning.data_ptr(), r.img.data_ptr()
            a.dL_dout_color, a.dL_dout_depth, a.dL_dout_alpha = _ptr(grad_color), _ptr(grad_depth), _ptr(grad_alpha)
            a.debug, a.grad_scratch = int(r.debug), _ptr(scratch)
            a.dL_dmeans2D, a.dL_dcolors, a.dL_dopacity, a.dL_dmeans3D = _ptr(g_means2D), _ptr(g_colors), _ptr(g_opacity), _ptr(g_means3D)
            a.dL_dcov3D, a.dL_dsh, a.dL_dscales, a.dL_drotations = _ptr(g_cov3D), _ptr(g_sh), _ptr(g_scales), _ptr(g_rot)

            def run():
                with torch.cuda.device(dev):
                    _lib.check(L.d2gs_gs3d_backward(C.byref(a), _stream_ptr(dev)), "d2gs_gs3d_backward")
            if r.debug:
                try:
                    run()
                except Exception as ex:
                    torch.save(cpu_deep_copy_tuple((means3D, radii, colors_precomp, scales, rotations, cov3Ds_precomp, grad_color,
                                                    grad_depth, grad_alpha, sh, alpha)), "snapshot_bw.dump")
                    print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                    raise ex
            else:
                run()
        # slots: means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings
        return (g_means3D, g_means2D, g_sh if sh is not None else None, g_colors if colors_precomp is not None else None,
                g_opacity, g_scales if scales is not None else None, g_rot if rotations is not None else None,
                g_cov3D if cov3Ds_precomp is not None else None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


def export_state(rctx: Optional[Gs3dContext] = None) -> dict:
    """Parity/debug: the per-Gaussian records, tile lists and per-pixel counters of a forward call, as torch tensors."""
    L = _lib.lib()
    r = rctx or LAST_CONTEXT
    dev = r.geom.device
    P, W, H, R = r.P, r.W, r.H, r.R
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out = dict(rec=torch.zeros((P, 12), dtype=torch.float32, device=dev), cov3D=torch.zeros((P, 6), dtype=torch.float32, device=dev),
               clamped=torch.zeros((P,), dtype=torch.uint8, device=dev), tiles_touched=torch.zeros((P,), dtype=torch.int32, device=dev),
               keys_sorted=torch.zeros((R,), dtype=torch.int64, device=dev), point_list=torch.zeros((R,), dtype=torch.int32, device=dev),
               ranges=torch.zeros((tiles, 2), dtype=torch.int32, device=dev), n_contrib=torch.zeros((H, W), dtype=torch.int32, device=dev))
    s = _lib.Gs3dState()
    for k, t in out.items():
        setattr(s, k, t.data_ptr())
    with torch.cuda.device(dev):
        _lib.check(L.d2gs_gs3d_export_state(P, W, H, R, r.geom.data_ptr(), r.binning.data_ptr(), r.img.data_ptr(), C.byref(s),
                                            _stream_ptr(dev)), "d2gs_gs3d_export_state")
    out["means2D"], out["depths"] = out["rec"][:, 0:2], out["rec"][:, 2]
    out["conic_opacity"], out["rgb"] = out["rec"][:, 4:8], out["rec"][:, 8:11]
    out["num_rendered"] = R
    return out
