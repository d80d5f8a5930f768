"""Autograd front-end of the B200 surfel rasterizer (calls libd2gs.so through the C ABI of include/d2gs.h).

Mirrors ``_RasterizeGaussians`` of the reference op (DSR/diff_surfel_rasterization/__init__.py:44-156): same
argument order, same outputs ``(color (3,H,W), radii (P) int32, allmap (8,H,W))``, same gradient slots, same
debug-snapshot behaviour.  Extension over the reference: ``sh`` may be given as two tensors (DC block + rest)
so the caller never concatenates 192 B per surfel per frame (gaussian_renderer/__init__.py:114,122).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from typing import Optional

import torch

from . import _lib

# (device index, P, W, H) -> last known num_rendered; sizes the binning workspace so forward is one C call
_R_HINT: dict = {}
# Deferred-count mode (include/d2gs.h "binning_capacity"), OPT-IN (set_deferred_count(True); always used while a CUDA graph
# is being captured): after `warmup` synchronous frames of a (device, P, W, H) combination the forward stops reading the
# instance count back and bins into max-seen-count * margin slots, so the host runs ahead of the device.  Counts arrive
# through pinned memory and are folded in by later calls; a frame that would overflow its slots renders NaN and the next
# rasterizer call raises (the following frames use the larger count).  The default is the reference's behaviour — one
# blocking 4-byte readback per forward (rasterizer_impl.cu:281-282) — which can never overflow, whatever the camera does.
_DEFERRED = {"on": False, "warmup": 4, "margin": 1.5}
_TRACK: "OrderedDict" = None          # hint key -> _CountTrack (LRU, see _lru_get)
# (device index, stream, P) -> zeroed (P,20) fp32 scratch the blend backward accumulates into (left zero by the library);
# keyed by stream so that backward passes running concurrently on two streams never share one
_GRAD_SCRATCH: "OrderedDict" = None
_LAUNCH_COUNT = {"forward": 0, "backward": 0}
_TRACK = OrderedDict()
_GRAD_SCRATCH = OrderedDict()
_TRACK_MAX, _SCRATCH_MAX = 64, 8


def _lru_get(cache: OrderedDict, key, make, limit: int, on_evict=None):
    """Least-recently-used lookup: densification changes P every few hundred iterations, so old (P, ...) entries age out
    one at a time instead of the whole history being dropped."""
    v = cache.get(key)
    if v is None:
        v = cache[key] = make()
        while len(cache) > limit:
            old, _ = cache.popitem(last=False)
            if on_evict is not None:
                on_evict(old)
    else:
        cache.move_to_end(key)
    return v


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _opt(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The reference encodes "not provided" as an empty tensor (DSR/.../__init__.py:198-208)."""
    if t is None or t.numel() == 0:
        return None
    return t


def _prep(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def _scratch_key(device: torch.device, P: int):
    return (device.index, torch.cuda.current_stream(device).cuda_stream, P)


def grad_scratch(device: torch.device, P: int) -> torch.Tensor:
    """Zeroed (P,20) accumulator of the blend backward for the CURRENT stream of `device` (kernels on one stream are
    ordered, so one buffer per (stream, P) is never shared by two backward passes in flight)."""
    return _lru_get(_GRAD_SCRATCH, _scratch_key(device, P),
                    lambda: torch.zeros((max(P, 1), 20), dtype=torch.float32, device=device), _SCRATCH_MAX)


def set_deferred_count(on: bool = True, warmup: int = 4, margin: float = 1.5) -> None:
    """Switch the host-synchronisation-free binning mode (default OFF = the reference's behaviour: one stream
    synchronisation per forward to read num_rendered, rasterizer_impl.cu:281-282).  Turn it on for loops whose cameras
    keep the instance count within ``margin`` of what was seen before (bench.py does; CUDA-graph capture implies it)."""
    _DEFERRED.update(on=bool(on), warmup=int(warmup), margin=float(margin))


class _CountTrack:
    __slots__ = ("max_R", "frames", "pending", "overflow", "captured", "spare")

    def __init__(self):
        self.max_R, self.frames, self.pending, self.overflow = 0, 0, [], None
        self.spare = []        # pinned count buffers set aside for captures (pinned allocation is illegal while capturing)
        self.captured = []     # (pinned {R, overflow}, capacity) of frames recorded into CUDA graphs: rewritten by every replay

    def observe(self, R: int) -> None:
        # slow decay so the capacity follows a scene that shrinks; growth is immediate
        self.max_R = max(int(R), int(self.max_R * 0.999))
        self.frames += 1

    def poll(self, key, block: bool = False) -> None:
        """Fold in the counts of earlier deferred frames whose copy has landed (oldest first; they complete in order)."""
        done = 0
        for ev, host, cap in self.pending:
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            R, ovf = host.tolist()
            R &= 0xffffffff
            self.observe(R)
            _R_HINT[key] = R
            if ovf:
                self.overflow = (R, cap)
            done += 1
        if done:
            del self.pending[:done]
        for host, cap in self.captured:      # values of the most recent completed replay (no event: a replay is not a call)
            R, ovf = host.tolist()
            R &= 0xffffffff
            if R:
                self.max_R = max(self.max_R, R)
                _R_HINT[key] = R
            if ovf:
                self.overflow = (R, cap)

    def raise_if_overflowed(self) -> None:
        if self.overflow is not None:
            R, cap = self.overflow
            self.overflow = None
            raise _lib.D2gsError(
                f"an earlier frame produced {R} (surfel, tile) instances but was binned into {cap} slots (deferred-count "
                "mode): its images were filled with NaN and its gradients are zero.  The capacity has been raised; repeat "
                "the step, or call d2gs_b200.raster.set_deferred_count(False) for views that change this abruptly.")


class RasterContext:
    """Opaque forward state kept for backward (the reference's geomBuffer / binningBuffer / imgBuffer).

    ``layout_R`` is the number of instance slots the binning workspace was laid out for (what the C ABI calls
    num_rendered in backward / export_state); ``num_rendered`` is the true instance count — in deferred-count mode
    reading it waits for the asynchronous copy of the count."""
    __slots__ = ("P", "D", "M", "W", "H", "layout_R", "_count", "_count_event", "_count_host", "geom", "binning", "img")

    @property
    def num_rendered(self) -> int:
        if self._count is None:
            if self._count_event is None:      # recorded into a CUDA graph: the last replay's count
                torch.cuda.synchronize(self.geom.device)
                return int(self._count_host[0]) & 0xffffffff
            self._count_event.synchronize()
            self._count = int(self._count_host[0]) & 0xffffffff
        return self._count


def raster_forward(bg, means3D, colors_precomp, opacities, scales, rotations, scale_modifier, transMat_precomp,
                   viewmatrix, projmatrix, tanfovx, tanfovy, image_height, image_width, sh, sh_rest, degree, campos,
                   prefiltered, debug, raw_params=False, d_means3D=None, d_scales=None, d_rotations=None):
    """Equivalent of ``_C.rasterize_gaussians`` (DSR/rasterize_points.cu:39-141).

    Returns (num_rendered, color, others, radii, ctx)."""
    L = _lib.lib()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if scales is not None and (scales.dim() != 2 or scales.shape[1] != 2):
        raise RuntimeError("scales must have dimensions (num_points, 2)")
    if rotations is not None and (rotations.dim() != 2 or rotations.shape[1] != 4):
        raise RuntimeError("rotations must have dimensions (num_points, 4)")
    dev = means3D.device
    P, H, W = int(means3D.shape[0]), int(image_height), int(image_width)
    M = 0
    if sh is not None:
        M = int(sh.shape[1]) + (int(sh_rest.shape[1]) if sh_rest is not None else 0)
    f32 = dict(dtype=torch.float32, device=dev)
    color = torch.empty((3, H, W), **f32)
    others = torch.empty((8, H, W), **f32)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    ctx = RasterContext()
    ctx.P, ctx.D, ctx.M, ctx.W, ctx.H = P, int(degree), M, W, H
    gbytes, ibytes, _ = _lib.workspace_sizes(P, W, H, 0)
    ctx.geom = torch.empty((gbytes,), dtype=torch.uint8, device=dev)
    ctx.img = torch.empty((ibytes,), dtype=torch.uint8, device=dev)
    hint_key = (dev.index, P, W, H)
    track = _lru_get(_TRACK, hint_key, _CountTrack, _TRACK_MAX, on_evict=lambda k: _R_HINT.pop(k, None))
    # CUDA-graph capture: nothing may synchronise or query, so the frame is recorded in deferred-count mode with the
    # capacity the eager frames before it established (the count still lands in pinned memory on every replay:
    # check_deferred_counts()).
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing:
        if debug or P == 0 or track.max_R <= 0:
            raise _lib.D2gsError("CUDA-graph capture of the rasterizer needs debug=False and at least one eager frame of this "
                                 "(P, width, height) first: the binning capacity comes from observed instance counts")
        deferred = True
    else:
        track.poll(hint_key)
        track.raise_if_overflowed()
        if not track.spare:
            track.spare = [torch.zeros((2,), dtype=torch.int32).pin_memory() for _ in range(4)]
        deferred = bool(_DEFERRED["on"]) and P > 0 and not debug and track.frames >= _DEFERRED["warmup"] and track.max_R > 0
    if deferred:
        capacity = int(track.max_R * _DEFERRED["margin"]) + 4096
    else:
        capacity = int(_R_HINT.get(hint_key, 4 * P + 1024) * 1.25) + 1024
    _, _, bbytes = _lib.workspace_sizes(P, W, H, capacity)
    ctx.binning = torch.empty((bbytes,), dtype=torch.uint8, device=dev)

    num_rendered = C.c_int64(0)
    required = C.c_size_t(0)
    a = _lib.RasterFwdArgs()
    a.P, a.D, a.M, a.width, a.height = P, int(degree), M, W, H
    a.background = _ptr(bg); a.means3D = _ptr(means3D)
    a.shs = _ptr(sh); a.sh_rest = _ptr(sh_rest); a.colors_precomp = _ptr(colors_precomp)
    a.opacities = _ptr(opacities); a.scales = _ptr(scales); a.rotations = _ptr(rotations)
    a.transMat_precomp = _ptr(transMat_precomp); a.scale_modifier = float(scale_modifier)
    a.viewmatrix = _ptr(viewmatrix); a.projmatrix = _ptr(projmatrix); a.campos = _ptr(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.out_color, a.out_others, a.radii = color.data_ptr(), others.data_ptr(), radii.data_ptr()
    a.geom_buffer, a.geom_bytes = ctx.geom.data_ptr(), gbytes
    a.img_buffer, a.img_bytes = ctx.img.data_ptr(), ibytes
    a.binning_buffer, a.binning_bytes = ctx.binning.data_ptr(), bbytes
    a.resume = 0
    a.raw_params = int(bool(raw_params))
    a.d_means3D, a.d_scales, a.d_rotations = _ptr(d_means3D), _ptr(d_scales), _ptr(d_rotations)
    a.num_rendered = C.pointer(num_rendered)
    a.binning_required = C.pointer(required)
    ctx._count, ctx._count_event, ctx._count_host = None, None, None
    if deferred:
        if capturing:
            if not track.spare:
                raise _lib.D2gsError("more than 4 rasterizer frames of one (P, width, height) captured without an eager frame in between")
            host = track.spare.pop()
        else:
            host = torch.empty((2,), dtype=torch.int32, pin_memory=True)   # torch's pinned allocator recycles these
        a.binning_capacity, a.num_rendered_async = capacity, host.data_ptr()
        ctx._count_host = host
    stream = _stream_ptr(dev)
    with torch.cuda.device(dev):
        rc = L.d2gs_raster_forward(C.byref(a), stream)
        if rc == _lib.D2GS_NEED_BINNING:
            bbytes = int(required.value * 1.1) + 4096
            ctx.binning = torch.empty((bbytes,), dtype=torch.uint8, device=dev)
            a.binning_buffer, a.binning_bytes, a.resume = ctx.binning.data_ptr(), bbytes, 1
            rc = L.d2gs_raster_forward(C.byref(a), stream)
        if deferred and rc == _lib.D2GS_OK and not capturing:
            ctx._count_event = torch.cuda.Event()
            ctx._count_event.record(torch.cuda.current_stream(dev))
    _lib.check(rc, "d2gs_raster_forward")
    if rc != _lib.D2GS_OK:
        raise _lib.D2gsError(f"d2gs_raster_forward returned {rc}")
    ctx.layout_R = int(num_rendered.value)
    if capturing:
        track.captured.append((host, capacity))
    elif deferred:
        track.pending.append((ctx._count_event, host, capacity))
    else:
        ctx._count = ctx.layout_R
        track.observe(ctx._count)
        _R_HINT[hint_key] = ctx._count
    _LAUNCH_COUNT["forward"] += 1
    return ctx.layout_R, color, others, radii, ctx


def raster_backward(bg, means3D, radii, colors_precomp, scales, rotations, scale_modifier, transMat_precomp, viewmatrix,
                    projmatrix, tanfovx, tanfovy, dL_dout_color, dL_dout_others, sh, sh_rest, degree, campos, ctx,
                    debug, want=None, raw_params=False, opacities=None, d_means3D=None, d_scales=None, d_rotations=None, out=None):
    """Equivalent of ``_C.rasterize_gaussians_backward`` (DSR/rasterize_points.cu:143-240).

    ``out``: optional dict of preallocated gradient tensors (e.g. slices of a gradient bucket); every element is written.

    Returns dict with dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dtransMat, dL_dsh, dL_dsh_rest,
    dL_dscales, dL_drotations (entries not requested through ``want`` are None)."""
    L = _lib.lib()
    dev = means3D.device
    P, M = ctx.P, ctx.M
    f32 = dict(dtype=torch.float32, device=dev)
    want = want or {}
    w = lambda k: want.get(k, True)
    g = dict.fromkeys(("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dtransMat", "dL_dsh",
                       "dL_dsh_rest", "dL_dscales", "dL_drotations", "dL_dscales_raw"))
    out = out or {}

    def buf(k, shape):
        t = out.get(k)
        if t is not None:
            assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == math.prod(shape), k
            return t.view(shape)
        return torch.empty(shape, **f32)

    if raw_params:
        g["dL_dscales_raw"] = buf("dL_dscales_raw", (P, 2))
    if w("dL_dmeans2D"): g["dL_dmeans2D"] = buf("dL_dmeans2D", (P, 3))
    if w("dL_dcolors"): g["dL_dcolors"] = buf("dL_dcolors", (P, 3))
    if w("dL_dopacity"): g["dL_dopacity"] = buf("dL_dopacity", (P, 1))
    if w("dL_dmeans3D"): g["dL_dmeans3D"] = buf("dL_dmeans3D", (P, 3))
    if w("dL_dtransMat"): g["dL_dtransMat"] = buf("dL_dtransMat", (P, 9))
    if sh is not None and w("dL_dsh"):
        if sh_rest is not None:
            g["dL_dsh"] = buf("dL_dsh", (P, 1, 3))
            g["dL_dsh_rest"] = buf("dL_dsh_rest", (P, M - 1, 3))
        else:
            g["dL_dsh"] = buf("dL_dsh", (P, M, 3))
    elif w("dL_dsh"):
        g["dL_dsh"] = torch.zeros((P, 0, 3), **f32)
    if w("dL_dscales"): g["dL_dscales"] = buf("dL_dscales", (P, 2))
    if w("dL_drotations"): g["dL_drotations"] = buf("dL_drotations", (P, 4))
    if P == 0:
        return g
    if scales is None:   # transMat_precomp path: no surfel-frame gradients
        for k in ("dL_dscales", "dL_drotations"):
            if g[k] is not None: g[k].zero_()
    a = _lib.RasterBwdArgs()
    a.P, a.D, a.M, a.width, a.height = P, int(degree), M, ctx.W, ctx.H
    a.num_rendered = ctx.layout_R
    a.background = _ptr(bg); a.means3D = _ptr(means3D); a.shs = _ptr(sh); a.sh_rest = _ptr(sh_rest)
    a.colors_precomp = _ptr(colors_precomp); a.scales = _ptr(scales); a.rotations = _ptr(rotations)
    a.transMat_precomp = _ptr(transMat_precomp); a.scale_modifier = float(scale_modifier)
    a.viewmatrix = _ptr(viewmatrix); a.projmatrix = _ptr(projmatrix); a.campos = _ptr(campos)
    a.tan_fovx, a.tan_fovy = float(tanfovx), float(tanfovy)
    a.radii = radii.data_ptr()
    a.geom_buffer, a.binning_buffer, a.img_buffer = ctx.geom.data_ptr(), ctx.binning.data_ptr(), ctx.img.data_ptr()
    a.dL_dout_color, a.dL_dout_others = dL_dout_color.data_ptr(), dL_dout_others.data_ptr()
    a.debug = int(bool(debug))
    a.grad_scratch = grad_scratch(dev, P).data_ptr()
    a.dL_dmeans2D = _ptr(g["dL_dmeans2D"]); a.dL_dcolors = _ptr(g["dL_dcolors"]); a.dL_dopacity = _ptr(g["dL_dopacity"])
    a.dL_dmeans3D = _ptr(g["dL_dmeans3D"]); a.dL_dtransMat = _ptr(g["dL_dtransMat"])
    a.dL_dsh = _ptr(g["dL_dsh"]) if sh is not None else None
    a.dL_dsh_rest = _ptr(g["dL_dsh_rest"])
    a.dL_dscales = _ptr(g["dL_dscales"]) if scales is not None else None
    a.dL_drotations = _ptr(g["dL_drotations"]) if scales is not None else None
    a.raw_params = int(bool(raw_params))
    a.opacities, a.d_means3D, a.d_scales, a.d_rotations = _ptr(opacities), _ptr(d_means3D), _ptr(d_scales), _ptr(d_rotations)
    a.dL_dscales_raw = _ptr(g["dL_dscales_raw"])
    with torch.cuda.device(dev):
        rc = L.d2gs_raster_backward(C.byref(a), _stream_ptr(dev))
    if rc != 0:
        _GRAD_SCRATCH.pop(_scratch_key(dev, P), None)   # scratch may be dirty: never reuse it
    _lib.check(rc, "d2gs_raster_backward")
    _LAUNCH_COUNT["backward"] += 1
    return g


class _RasterizeSurfels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, sh_rest, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        rs = raster_settings
        sh_, shr_ = _prep(_opt(sh), "sh"), _prep(_opt(sh_rest), "sh_rest")
        col_ = _prep(_opt(colors_precomp), "colors")
        sc_, rot_ = _prep(_opt(scales), "scales"), _prep(_opt(rotations), "rotations")
        cov_ = _prep(_opt(cov3Ds_precomp), "transMat_precomp")
        m3 = _prep(means3D, "means3D")
        op_ = _prep(opacities, "opacity")
        bg = _prep(rs.bg, "background")
        view, proj, campos = _prep(rs.viewmatrix, "viewmatrix"), _prep(rs.projmatrix, "projmatrix"), _prep(rs.campos, "campos")
        args = (bg, m3, col_, op_, sc_, rot_, rs.scale_modifier, cov_, view, proj, rs.tanfovx, rs.tanfovy,
                rs.image_height, rs.image_width, sh_, shr_, rs.sh_degree, campos, rs.prefiltered, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)   # copy before anything can be corrupted
            try:
                num_rendered, color, others, radii, rctx = raster_forward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, others, radii, rctx = raster_forward(*args)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered   # slot count of the binning layout (== the instance count in synchronous mode)
        ctx.rctx = rctx
        ctx.prepped = (bg, view, proj, campos)
        ctx.save_for_backward(col_, m3, sc_, rot_, cov_, radii, sh_, shr_)
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        col_, m3, sc_, rot_, cov_, radii, sh_, shr_ = ctx.saved_tensors
        bg, view, proj, campos = ctx.prepped
        rctx = ctx.rctx
        dev = m3.device
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, rctx.H, rctx.W), dtype=torch.float32, device=dev)
        if grad_depth is None:
            grad_depth = torch.zeros((8, rctx.H, rctx.W), dtype=torch.float32, device=dev)
        grad_out_color = grad_out_color.float().contiguous()
        grad_depth = grad_depth.float().contiguous()
        need = ctx.needs_input_grad
        want = {"dL_dmeans3D": need[0], "dL_dmeans2D": need[1], "dL_dsh": need[2] or need[3],
                "dL_dcolors": need[4] and col_ is not None, "dL_dopacity": need[5], "dL_dscales": need[6] or need[7],
                "dL_drotations": need[6] or need[7], "dL_dtransMat": need[8] and cov_ is not None}
        args = (bg, m3, radii, col_, sc_, rot_, rs.scale_modifier, cov_, view, proj, rs.tanfovx, rs.tanfovy,
                grad_out_color, grad_depth, sh_, shr_, rs.sh_degree, campos, rctx, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args[:-2])
            try:
                g = raster_backward(*args, want=want)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            g = raster_backward(*args, want=want)
        grad_sh, grad_sh_rest = g["dL_dsh"], g["dL_dsh_rest"]
        if sh_ is None:
            grad_sh = None
        return (g["dL_dmeans3D"], g["dL_dmeans2D"], grad_sh, grad_sh_rest, g["dL_dcolors"], g["dL_dopacity"],
                g["dL_dscales"] if sc_ is not None else None, g["dL_drotations"] if rot_ is not None else None,
                g["dL_dtransMat"] if cov_ is not None else None, None)


class _RasterizeSurfelsRaw(torch.autograd.Function):
    """Rasterizer fused with the activations and deformation deltas of render() (raw-parameter mode of the C ABI).

    inputs: xyz, d_xyz, log-scales, d_scaling, raw quaternions, d_rotation, opacity logits (deltas may be None /
    python scalars 0.0), SH as (DC, rest).  Semantics == rasterize(xyz+d_xyz, exp(s)+d_s, normalize(q+d_q), sigmoid(o))."""

    @staticmethod
    def forward(ctx, xyz, d_xyz, scaling, d_scaling, rotation, d_rotation, opacity, means2D, sh, sh_rest, colors_precomp,
                raster_settings):
        rs = raster_settings
        t = lambda v, n: _prep(v, n) if torch.is_tensor(v) else None
        xyz_, sc_, rot_, op_ = _prep(xyz, "xyz"), _prep(scaling, "scaling"), _prep(rotation, "rotation"), _prep(opacity, "opacity")
        dx_, ds_, dr_ = t(d_xyz, "d_xyz"), t(d_scaling, "d_scaling"), t(d_rotation, "d_rotation")
        for v, like, n in ((dx_, xyz_, "d_xyz"), (ds_, sc_, "d_scaling"), (dr_, rot_, "d_rotation")):
            if v is not None and v.shape != like.shape:
                raise RuntimeError(f"{n} must have the shape of the parameter it offsets, got {tuple(v.shape)}")
        sh_, shr_, col_ = _prep(_opt(sh), "sh"), _prep(_opt(sh_rest), "sh_rest"), _prep(_opt(colors_precomp), "colors")
        bg = _prep(rs.bg, "background")
        view, proj, campos = _prep(rs.viewmatrix, "viewmatrix"), _prep(rs.projmatrix, "projmatrix"), _prep(rs.campos, "campos")
        num_rendered, color, others, radii, rctx = raster_forward(
            bg, xyz_, col_, op_, sc_, rot_, rs.scale_modifier, None, view, proj, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, sh_, shr_, rs.sh_degree, campos, rs.prefiltered, rs.debug, raw_params=True, d_means3D=dx_,
            d_scales=ds_, d_rotations=dr_)
        ctx.raster_settings, ctx.rctx, ctx.prepped = rs, rctx, (bg, view, proj, campos)
        ctx.has_delta = (dx_ is not None, ds_ is not None, dr_ is not None)
        ctx.save_for_backward(xyz_, sc_, rot_, op_, dx_, ds_, dr_, radii, sh_, shr_, col_)
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        xyz_, sc_, rot_, op_, dx_, ds_, dr_, radii, sh_, shr_, col_ = ctx.saved_tensors
        bg, view, proj, campos = ctx.prepped
        rctx = ctx.rctx
        dev = xyz_.device
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, rctx.H, rctx.W), dtype=torch.float32, device=dev)
        if grad_depth is None:
            grad_depth = torch.zeros((8, rctx.H, rctx.W), dtype=torch.float32, device=dev)
        want = {"dL_dtransMat": False, "dL_dcolors": col_ is not None}
        # parameter gradients are written straight into the gradient bucket when one owns the parameter (dist.claim)
        from .dist import claim
        out = {"dL_dmeans3D": claim(xyz_, False), "dL_dscales_raw": claim(sc_, False), "dL_drotations": claim(rot_, False),
               "dL_dopacity": claim(op_, False), "dL_dsh": claim(sh_, False), "dL_dsh_rest": claim(shr_, False)}
        g = raster_backward(bg, xyz_, radii, col_, sc_, rot_, rs.scale_modifier, None, view, proj, rs.tanfovx, rs.tanfovy,
                            grad_out_color.float().contiguous(), grad_depth.float().contiguous(), sh_, shr_, rs.sh_degree,
                            campos, rctx, rs.debug, want=want, raw_params=True, opacities=op_, d_means3D=dx_, d_scales=ds_,
                            d_rotations=dr_, out=out)
        from .dist import grads_ready, reduces_early
        hd = ctx.has_delta

        def for_delta(grad, param):
            # dL/d(delta) == dL/d(parameter).  Normally a second tensor object on the same memory (autograd adopts the
            # first as .grad); but when the parameter's bucket slot is all-reduced EARLY — in place, on NCCL's stream, from
            # grads_ready() below — the deformation backward would read it while other ranks' values are being summed in:
            # it gets a private copy, taken before the collective is launched.
            return grad.clone() if reduces_early(param) else grad.view_as(grad)
        g_dxyz = for_delta(g["dL_dmeans3D"], xyz_) if hd[0] else None
        g_drot = for_delta(g["dL_drotations"], rot_) if hd[2] else None
        grads_ready("raster")      # the surfel-table gradients are final: buckets may start reducing them now
        return (g["dL_dmeans3D"], g_dxyz, g["dL_dscales_raw"], g["dL_dscales"] if hd[1] else None,
                g["dL_drotations"], g_drot, g["dL_dopacity"], g["dL_dmeans2D"],
                g["dL_dsh"] if sh_ is not None else None, g["dL_dsh_rest"], g["dL_dcolors"], None)


def rasterize_surfels_raw(xyz, d_xyz, scaling, d_scaling, rotation, d_rotation, opacity, means2D, sh, sh_rest,
                          colors_precomp, raster_settings):
    return _RasterizeSurfelsRaw.apply(xyz, d_xyz, scaling, d_scaling, rotation, d_rotation, opacity, means2D, sh, sh_rest,
                                      colors_precomp, raster_settings)


def rasterize_surfels(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                      raster_settings, sh_rest=None):
    """Functional entry point; with ``sh_rest`` the SH coefficients are passed as (P,1,3) + (P,M-1,3)."""
    return _RasterizeSurfels.apply(means3D, means2D, sh, sh_rest, colors_precomp, opacities, scales, rotations,
                                   cov3Ds_precomp, raster_settings)


def mark_visible(positions, viewmatrix, projmatrix):
    """Equivalent of ``_C.mark_visible`` (DSR/rasterize_points.cu:242-261)."""
    L = _lib.lib()
    pos = _prep(positions, "means3D")
    view, proj = _prep(viewmatrix, "viewmatrix"), _prep(projmatrix, "projmatrix")
    P = int(pos.shape[0])
    present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
    if P:
        with torch.cuda.device(pos.device):
            _lib.check(L.d2gs_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(), present.data_ptr(),
                                           _stream_ptr(pos.device)), "d2gs_mark_visible")
    return present


def export_state(rctx: RasterContext) -> dict:
    """Parity/debug: the intermediates the reference hides in its three byte buffers, as torch tensors."""
    L = _lib.lib()
    dev = rctx.geom.device
    P, W, H, R = rctx.P, rctx.W, rctx.H, rctx.layout_R
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(means2D=torch.zeros((P, 2), **f32), depths=torch.zeros((P,), **f32), transMat=torch.zeros((P, 9), **f32),
               normal_opacity=torch.zeros((P, 4), **f32), rgb=torch.zeros((P, 3), **f32),
               clamped=torch.zeros((P, 3), dtype=torch.uint8, device=dev),
               tiles_touched=torch.zeros((P,), dtype=torch.int32, device=dev),
               point_offsets=torch.zeros((P,), dtype=torch.int32, device=dev),
               keys_unsorted=torch.zeros((R,), dtype=torch.int64, device=dev),
               values_unsorted=torch.zeros((R,), dtype=torch.int32, device=dev),
               keys_sorted=torch.zeros((R,), dtype=torch.int64, device=dev),
               point_list=torch.zeros((R,), dtype=torch.int32, device=dev),
               ranges=torch.zeros((tiles, 2), dtype=torch.int32, device=dev),
               final_T=torch.zeros((3, H, W), **f32), n_contrib=torch.zeros((2, H, W), dtype=torch.int32, device=dev))
    st = _lib.RasterState()
    for k, v in out.items():
        setattr(st, k, v.data_ptr() if v.numel() else None)
    with torch.cuda.device(dev):
        _lib.check(L.d2gs_raster_export_state(P, W, H, R, rctx.geom.data_ptr(), rctx.binning.data_ptr(),
                                              rctx.img.data_ptr(), C.byref(st), _stream_ptr(dev)),
                   "d2gs_raster_export_state")
    n = rctx.num_rendered        # deferred-count mode lays the workspace out for more slots than there are instances
    if n < R:
        for k in ("keys_unsorted", "values_unsorted", "keys_sorted", "point_list"):
            out[k] = out[k][:n]
    return out


def check_deferred_counts(device=None, wait: bool = True) -> None:
    """Fold in every outstanding asynchronous instance count (eager deferred frames and CUDA-graph replays) and raise
    if one of those frames overflowed its binning capacity.  For loops that only replay graphs — where no later
    rasterizer call would report it — call this after synchronising; an overflowed frame is NaN either way."""
    want = None if device is None else torch.device(device).index        # "cuda" without an index = every device
    for key, track in list(_TRACK.items()):
        dev_index = key[1] if key[0] == "gs3d" else key[0]       # the gs3d side path shares the tracker under ("gs3d", device, P, W, H)
        if want is not None and dev_index != want:
            continue
        track.poll(key, block=wait)
        track.raise_if_overflowed()


def last_num_rendered(device, P: int, W: int, H: int, wait: bool = True) -> int:
    """Most recent instance count seen for this (device, P, W, H) — waits for pending asynchronous counts by default."""
    key = (torch.device(device).index, P, W, H)
    track = _TRACK.get(key)
    if track is not None:
        track.poll(key, block=wait)
    return int(_R_HINT.get(key, 0))


def launch_counts() -> dict:
    return dict(_LAUNCH_COUNT)
