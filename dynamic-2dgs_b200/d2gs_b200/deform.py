"""Node-controlled deformation, B200 path: the time-conditioned MLP evaluated at the M control nodes, then the
fused CUDA KNN + radial-basis weighting + blend onto the P surfels (libd2gs.so: d2gs_deform_forward/backward).

Mirrors the reference interface (paths relative to /root/reference):
  * ``DeformNetwork``      utils/time_utils.py:310-458   same constructor, parameter names and output dict
  * ``ControlNodeWarp``    utils/time_utils.py:770-1233  same constructor, ``forward(x, t, feature, motion_mask, ...)``,
                           ``cal_nn_weight``, ``node_deform``, ``query_network``, ``expand_time``, state-dict keys
                           (``nodes``, ``_node_radius``, ``_node_weight``, ``network.*``, buffer ``inited``)
  * ``DeformModel``        scene/deform_model.py:13-72   ``step(xyz, time_emb, iteration=0, **kwargs)``
Two ways to use it:
  * stand-alone (no reference tree): the classes above, with plain-torch versions of the regularisers
    (``arap_loss`` / ``elastic_loss`` / ``acc_loss``) and of ``cal_nn_weight`` for the calls outside the fused kernel;
  * ``install_into_reference()`` (reference tree importable): SUBCLASSES of the reference's own ``ControlNodeWarp`` /
    ``DeformNetwork`` that override only ``forward`` / ``cal_nn_weight`` for the configurations the CUDA kernels cover and
    inherit everything else (``as_gaussians``, ``densify``, ``arap_loss``, editing branches, ``gs_*`` state) — registered
    in ``scene.deform_model.model_dict`` so train_gui.py runs unchanged.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Any, Mapping, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


# --------------------------------------------------------------------------------------------------------------
# fused KNN + weights + blend
# --------------------------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


def processing_order(xyz: torch.Tensor) -> torch.Tensor:
    """int32 permutation that sorts the surfel centres along a Morton curve (libd2gs.so: d2gs_deform_order).  Passed to
    the blend kernels it makes the 32 surfels of a warp spatial neighbours; results do not depend on it."""
    L = _lib.lib()
    x = _f32(xyz)
    P = int(x.shape[0])
    order = torch.empty((P,), dtype=torch.int32, device=x.device)
    if P:
        nbytes = C.c_size_t()
        _lib.check(L.d2gs_deform_order_workspace(P, C.byref(nbytes)), "d2gs_deform_order_workspace")
        ws = torch.empty((nbytes.value,), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(L.d2gs_deform_order(P, x.data_ptr(), order.data_ptr(), ws.data_ptr(), nbytes.value, _stream(x.device)),
                       "d2gs_deform_order")
    return order


def _blend_forward(ctx, xyz, feature, nodes, node_radius_log, node_weight_logit, attr_ptrs, attr_stride, motion_mask, K, hyper_dim,
                   order=None):
    """Shared forward of the two autograd Functions below.  attr_ptrs = device pointers of (trans, rot, scale, local_rot)."""
    L = _lib.lib()
    dev = xyz.device
    if not xyz.is_cuda:
        raise RuntimeError("xyz must be a CUDA tensor (the deformation blend has no CPU path)")
    xyz_, feat_, nodes_ = _f32(xyz), _f32(feature), _f32(nodes)
    rad_, wl_ = _f32(node_radius_log), _f32(node_weight_logit)
    mask_ = None
    if torch.is_tensor(motion_mask):
        mask_ = _f32(motion_mask).reshape(-1)
        if mask_.numel() == 1:
            mask_ = mask_.expand(xyz_.shape[0]).contiguous()
    P, M = int(xyz_.shape[0]), int(nodes_.shape[0])
    use_hyper = hyper_dim > 0 and feat_ is not None
    a = _lib.DeformFwdArgs()
    a.P, a.M, a.K, a.hyper_dim = P, M, int(K), int(nodes_.shape[1] - 3)
    a.xyz = _p(xyz_)
    a.feature = _p(feat_) if use_hyper else None
    a.feature_stride = int(feat_.shape[1]) if use_hyper else 0
    a.nodes, a.node_radius_log, a.node_weight_logit = _p(nodes_), _p(rad_), _p(wl_)
    a.node_trans, a.node_rot, a.node_scale, a.node_local_rot = attr_ptrs
    a.node_attr_stride = attr_stride
    a.motion_mask = _p(mask_)
    if order is not None:
        if order.dtype != torch.int32 or order.numel() != P or not order.is_contiguous() or order.device != dev:
            raise ValueError("order must be a contiguous int32 permutation of 0..P-1 on the same device")
        a.order = order.data_ptr()
    ctx.order = order
    nn_idx = torch.empty((P, K), dtype=torch.int64, device=dev)
    nn_dist = torch.empty((P, K), dtype=torch.float32, device=dev)
    nn_weight = torch.empty((P, K), dtype=torch.float32, device=dev)
    d_xyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
    d_rot = torch.empty((P, 4), dtype=torch.float32, device=dev)
    d_scale = torch.empty((P, 2), dtype=torch.float32, device=dev)
    a.nn_idx, a.nn_dist, a.nn_weight = _p(nn_idx), _p(nn_dist), _p(nn_weight)
    a.d_xyz, a.d_rotation, a.d_scaling = _p(d_xyz), _p(d_rot), _p(d_scale)
    with torch.cuda.device(dev):
        _lib.check(L.d2gs_deform_forward(C.byref(a), _stream(dev)), "d2gs_deform_forward")
    ctx.K, ctx.use_hyper = int(K), use_hyper
    ctx.mask_shape = motion_mask.shape if torch.is_tensor(motion_mask) else None
    ctx.mark_non_differentiable(nn_idx)
    return (xyz_, feat_, nodes_, rad_, wl_, mask_, nn_idx, nn_dist, nn_weight), (d_xyz, d_rot, d_scale, nn_weight, nn_dist, nn_idx)


def _blend_backward(ctx, common, attr_ptrs, d_attr_ptrs, attr_stride, g_xyz, g_rot, g_scale):
    """Shared backward.  Node-level outputs are accumulated with atomics: d_attr_ptrs must point at zeroed memory.
    Parameter gradients go straight into the gradient bucket when one has claimed the parameter (dist.claim)."""
    from . import dist as _dist
    L = _lib.lib()
    xyz_, feat_, nodes_, rad_, wl_, mask_, nn_idx, nn_dist, nn_weight = common
    dev = xyz_.device
    P, M, K = int(xyz_.shape[0]), int(nodes_.shape[0]), ctx.K
    zeros = lambda shape: torch.zeros(shape, dtype=torch.float32, device=dev)
    g_xyz = zeros((P, 3)) if g_xyz is None else g_xyz.float().contiguous()
    g_rot = zeros((P, 4)) if g_rot is None else g_rot.float().contiguous()
    g_scale = zeros((P, 2)) if g_scale is None else g_scale.float().contiguous()
    n_nodes, n_wl = nodes_.numel(), (wl_.numel() if wl_ is not None else 0)
    d_nodes, d_rad = _dist.claim(nodes_, zeroed=True), _dist.claim(rad_, zeroed=True)
    d_wl = _dist.claim(wl_, zeroed=True) if wl_ is not None else None
    if d_nodes is None or d_rad is None or (wl_ is not None and d_wl is None):
        flat = zeros((n_nodes + M + n_wl,))          # one fill for the three accumulated parameter gradients
        d_nodes = flat[:n_nodes].view_as(nodes_) if d_nodes is None else d_nodes
        d_rad = flat[n_nodes:n_nodes + M].view_as(rad_) if d_rad is None else d_rad
        if wl_ is not None and d_wl is None:
            d_wl = flat[n_nodes + M:].view_as(wl_)
    want_feat = feat_ is not None and ctx.use_hyper
    d_feat = None
    if want_feat:
        d_feat = _dist.claim(feat_, zeroed=False)
        if d_feat is None:
            d_feat = torch.empty_like(feat_)
    d_mask = torch.empty((P,), dtype=torch.float32, device=dev) if mask_ is not None else None
    a = _lib.DeformBwdArgs()
    a.P, a.M, a.K, a.hyper_dim = P, M, K, int(nodes_.shape[1] - 3)
    a.xyz = _p(xyz_)
    a.feature = _p(feat_) if ctx.use_hyper else None
    a.feature_stride = int(feat_.shape[1]) if ctx.use_hyper else 0
    a.nodes, a.node_radius_log, a.node_weight_logit = _p(nodes_), _p(rad_), _p(wl_)
    a.node_trans, a.node_rot, a.node_scale, a.node_local_rot = attr_ptrs
    a.node_attr_stride = attr_stride
    a.motion_mask = _p(mask_)
    if ctx.order is not None:
        a.order = ctx.order.data_ptr()
    a.nn_idx, a.nn_dist, a.nn_weight = _p(nn_idx), _p(nn_dist), _p(nn_weight)
    a.dL_d_xyz, a.dL_d_rotation, a.dL_d_scaling = _p(g_xyz), _p(g_rot), _p(g_scale)
    a.dL_dnode_trans, a.dL_dnode_rot, a.dL_dnode_scale, a.dL_dnode_local_rot = d_attr_ptrs
    a.dL_dnodes, a.dL_dnode_radius_log, a.dL_dnode_weight_logit = _p(d_nodes), _p(d_rad), _p(d_wl)
    a.dL_dfeature, a.dL_dmotion_mask = _p(d_feat), _p(d_mask)
    with torch.cuda.device(dev):
        _lib.check(L.d2gs_deform_backward(C.byref(a), _stream(dev)), "d2gs_deform_backward")
    _dist.grads_ready("deform")     # hyper-coordinate table and node geometry are final: a bucket may start reducing them
    if d_mask is not None and ctx.mask_shape is not None:
        d_mask = d_mask.sum().reshape(ctx.mask_shape) if math.prod(ctx.mask_shape) == 1 else d_mask.reshape(ctx.mask_shape)
    return d_feat, d_nodes, d_rad, d_wl, d_mask


class _NodeBlend(torch.autograd.Function):
    """(xyz, feature, nodes, log-radius, weight-logit, node outputs, mask) -> (d_xyz, d_rotation, d_scaling)."""

    @staticmethod
    def forward(ctx, xyz, feature, nodes, node_radius_log, node_weight_logit, node_trans, node_rot, node_scale,
                node_local_rot, motion_mask, K, hyper_dim, order=None):
        tr_, rt_, sc_, lr_ = _f32(node_trans), _f32(node_rot), _f32(node_scale), _f32(node_local_rot)
        common, outs = _blend_forward(ctx, xyz, feature, nodes, node_radius_log, node_weight_logit,
                                      (_p(tr_), _p(rt_), _p(sc_), _p(lr_)), 0, motion_mask, K, hyper_dim, order)
        ctx.has = [t is not None for t in common]
        ctx.save_for_backward(tr_, rt_, sc_, lr_, *[t for t in common if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_xyz, g_rot, g_scale, g_w, g_d, g_i):
        tr_, rt_, sc_, lr_, *rest = ctx.saved_tensors
        it = iter(rest)
        common = tuple(next(it) if h else None for h in ctx.has)
        M = int(tr_.shape[0])
        ncol = 9 + (4 if lr_ is not None else 0)
        flat = torch.zeros((M * ncol,), dtype=torch.float32, device=tr_.device)   # one fill for all node-level gradients
        d_trans, d_rot = flat[:3 * M].view(M, 3), flat[3 * M:7 * M].view(M, 4)
        d_scale = flat[7 * M:9 * M].view(M, 2)
        d_lr = flat[9 * M:].view(M, 4) if lr_ is not None else None
        d_feat, d_nodes, d_rad, d_wl, d_mask = _blend_backward(
            ctx, common, (_p(tr_), _p(rt_), _p(sc_), _p(lr_)), (_p(d_trans), _p(d_rot), _p(d_scale), _p(d_lr)), 0,
            g_xyz, g_rot, g_scale)
        # xyz is detached by the reference (time_utils.py:1136), node xyz columns too (:947,1151)
        return (None, d_feat, d_nodes, d_rad, d_wl, d_trans, d_rot, d_scale, d_lr, d_mask, None, None, None)


class _NodeBlendPacked(torch.autograd.Function):
    """Same op, with the per-node MLP outputs given as ONE (M, S) matrix (the fused MLP's head output) and the column of
    each attribute: no slicing copies forward, one (M, S) gradient matrix backward.  cols = (trans, rot, scale, local_rot | -1)."""

    @staticmethod
    def forward(ctx, xyz, feature, nodes, node_radius_log, node_weight_logit, attrs, motion_mask, K, hyper_dim, cols, order=None):
        at_ = _f32(attrs)
        S = int(at_.shape[1])
        base = at_.data_ptr()
        ptrs = tuple((base + 4 * c) if c >= 0 else None for c in cols)
        common, outs = _blend_forward(ctx, xyz, feature, nodes, node_radius_log, node_weight_logit, ptrs, S, motion_mask, K, hyper_dim,
                                      order)
        ctx.cols = tuple(cols)
        ctx.has = [t is not None for t in common]
        ctx.save_for_backward(at_, *[t for t in common if t is not None])
        return outs

    @staticmethod
    def backward(ctx, g_xyz, g_rot, g_scale, g_w, g_d, g_i):
        at_, *rest = ctx.saved_tensors
        it = iter(rest)
        common = tuple(next(it) if h else None for h in ctx.has)
        S = int(at_.shape[1])
        d_attrs = torch.zeros_like(at_)
        ptrs = tuple((at_.data_ptr() + 4 * c) if c >= 0 else None for c in ctx.cols)
        d_ptrs = tuple((d_attrs.data_ptr() + 4 * c) if c >= 0 else None for c in ctx.cols)
        d_feat, d_nodes, d_rad, d_wl, d_mask = _blend_backward(ctx, common, ptrs, d_ptrs, S, g_xyz, g_rot, g_scale)
        return (None, d_feat, d_nodes, d_rad, d_wl, d_attrs, d_mask, None, None, None, None)


def node_blend_packed(xyz, feature, nodes, node_radius_log, node_weight_logit, attrs, cols, motion_mask, K: int, hyper_dim: int,
                      order=None):
    d_xyz, d_rot, d_scale, w, d, i = _NodeBlendPacked.apply(xyz, feature, nodes, node_radius_log, node_weight_logit, attrs,
                                                            motion_mask, K, hyper_dim, tuple(cols), order)
    return {"d_xyz": d_xyz, "d_rotation": d_rot, "d_scaling": d_scale, "nn_weight": w, "nn_dist": d, "nn_index": i, "nn_idx": i}


def node_blend(xyz, feature, nodes, node_radius_log, node_weight_logit, node_trans, node_rot, node_scale,
               node_local_rot, motion_mask, K: int, hyper_dim: int, order=None):
    """Returns dict(d_xyz, d_rotation, d_scaling, nn_weight, nn_dist, nn_idx).  ``order``: see processing_order()."""
    d_xyz, d_rot, d_scale, w, d, i = _NodeBlend.apply(xyz, feature, nodes, node_radius_log, node_weight_logit, node_trans,
                                                      node_rot, node_scale, node_local_rot, motion_mask, K, hyper_dim, order)
    return {"d_xyz": d_xyz, "d_rotation": d_rot, "d_scaling": d_scale, "nn_weight": w, "nn_dist": d, "nn_idx": i}


# --------------------------------------------------------------------------------------------------------------
# fused MLP (libd2gs.so: d2gs_mlp_forward/backward)
# --------------------------------------------------------------------------------------------------------------
class _FusedMLP(torch.autograd.Function):
    """(x, t, is_blender, num_out, *weights) -> (out (rows,num_out), hidden (rows,256)).

    weights = [timenet.0.w, timenet.0.b, timenet.2.w, timenet.2.b] (is_blender only) + 8 x (linear.w, linear.b)
              + [heads_w (num_out,256), heads_b (num_out)]."""

    @staticmethod
    def forward(ctx, x, t, is_blender, num_out, *weights):
        L = _lib.lib()
        dev = x.device
        x_ = x.detach().float().contiguous()
        t_ = t.detach().float()
        rows = int(x_.shape[0])
        t2 = t_.reshape(rows, -1) if t_.numel() == rows else t_.expand(rows, 1)
        t_stride = int(t2.stride(0))
        if t_stride not in (0, 1):
            t2 = t2.contiguous(); t_stride = 1
        ws_bytes = C.c_size_t(0)
        _lib.check(L.d2gs_mlp_workspace(rows, int(is_blender), int(num_out), C.byref(ws_bytes)), "d2gs_mlp_workspace")
        ws = torch.empty((ws_bytes.value,), dtype=torch.uint8, device=dev)
        out = torch.empty((rows, num_out), dtype=torch.float32, device=dev)
        w = [p_.detach().float().contiguous() for p_ in weights]
        a = _FusedMLP._args(rows, is_blender, num_out, x_, t2, t_stride, w, ws)
        a.out = out.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_mlp_forward(C.byref(a), _stream(dev)), "d2gs_mlp_forward")
            hp = L.d2gs_mlp_hidden(rows, int(is_blender), int(num_out), ws.data_ptr())
        off = hp - ws.data_ptr()
        hidden = ws[off: off + rows * 256 * 4].view(torch.float32).view(rows, 256)
        ctx.cfg = (rows, bool(is_blender), int(num_out), t_stride)
        ctx.save_for_backward(x_, t2, ws, *w)
        ctx.mark_non_differentiable(hidden)
        return out, hidden

    @staticmethod
    def _args(rows, is_blender, num_out, x_, t2, t_stride, w, ws):
        a = _lib.MlpArgs()
        a.rows, a.is_blender, a.num_out = rows, int(is_blender), int(num_out)
        a.x, a.t, a.t_stride = x_.data_ptr(), t2.data_ptr(), t_stride
        i = 0
        if is_blender:
            a.timenet0_w, a.timenet0_b, a.timenet2_w, a.timenet2_b = (w[j].data_ptr() for j in range(4))
            i = 4
        for l in range(8):
            a.linear_w[l] = w[i + 2 * l].data_ptr()
            a.linear_b[l] = w[i + 2 * l + 1].data_ptr()
        a.heads_w, a.heads_b = w[i + 16].data_ptr(), w[i + 17].data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        return a

    @staticmethod
    def backward(ctx, g_out, g_hidden):
        L = _lib.lib()
        rows, is_blender, num_out, t_stride = ctx.cfg
        x_, t2, ws, *w = ctx.saved_tensors
        dev = x_.device
        g_out = torch.zeros((rows, num_out), dtype=torch.float32, device=dev) if g_out is None else g_out.float().contiguous()
        from .dist import claim
        grads = [claim(p_, False) for p_ in w]      # parameters owned by a gradient bucket are written in place
        grads = [torch.empty_like(p_) if g_ is None else g_ for p_, g_ in zip(w, grads)]
        a = _FusedMLP._args(rows, is_blender, num_out, x_, t2, t_stride, w, ws)
        a.g_out = g_out.data_ptr()
        i = 0
        if is_blender:
            a.g_timenet0_w, a.g_timenet0_b, a.g_timenet2_w, a.g_timenet2_b = (grads[j].data_ptr() for j in range(4))
            i = 4
        for l in range(8):
            a.g_linear_w[l] = grads[i + 2 * l].data_ptr()
            a.g_linear_b[l] = grads[i + 2 * l + 1].data_ptr()
        a.g_heads_w, a.g_heads_b = grads[i + 16].data_ptr(), grads[i + 17].data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.d2gs_mlp_backward(C.byref(a), _stream(dev)), "d2gs_mlp_backward")
        return (None, None, None, None, *grads)


# --------------------------------------------------------------------------------------------------------------
# embedder + MLP (same parameter names as the reference so deform.pth round-trips)
# --------------------------------------------------------------------------------------------------------------
class Embedder:
    """utils/time_utils.py:208-256: include_input, log-sampled frequencies 2^0..2^(L-1), [sin, cos] per band."""

    def __init__(self, multires: int, input_dims: int):
        self.multires, self.input_dims = multires, input_dims
        self.out_dim = input_dims * (1 + 2 * multires)
        self.freq_bands = 2.0 ** torch.linspace(0.0, multires - 1, steps=multires)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        outs = [x]
        for f in self.freq_bands.tolist():
            outs.append(torch.sin(x * f))
            outs.append(torch.cos(x * f))
        return torch.cat(outs, -1)


def get_embedder(multires, i=1):
    if i == -1:
        return nn.Identity(), 3
    e = Embedder(multires, i)
    return e, e.out_dim


class _FusedNetworkMixin:
    """Fast path of ``DeformNetwork.forward`` (the fused CUDA MLP); everything it does not cover goes to the next class
    in the MRO — the eager layers below, or the reference's own class (``bind_reference``)."""

    use_fused = True   # set False to force the eager torch layers (tests compare the two)

    def _fusable(self, x, t) -> bool:
        if not (self.use_fused and torch.is_tensor(x) and x.is_cuda and x.dim() == 2 and x.shape[1] == 3):
            return False
        if getattr(self, "pred_color", False) or getattr(self, "progressive_brand_time", False):
            return False
        xyz_ch = 63                                              # multires 10 on 3 coordinates, input included
        in0 = xyz_ch + (30 if self.is_blender else 21)           # timenet output | PE(t; 10 frequencies)
        lin = self.linear
        return (self.D == 8 and self.W == 256 and list(self.skips) == [4] and len(lin) == 8 and lin[0].in_features == in0
                and lin[5].in_features == 256 + in0 and t.shape[-1] == 1 and (t.numel() == x.shape[0] or t.numel() == 1))

    def _forward_fused(self, x, t):
        self._packed = None
        heads = [self.gaussian_warp, self.gaussian_scaling, self.gaussian_rotation]
        if self.local_frame:
            heads.append(self.local_rotation)
        if self.pred_opacity:
            heads.append(self.gaussian_opacity)
        hw = torch.cat([h.weight for h in heads], 0)
        hb = torch.cat([h.bias for h in heads], 0)
        ws = []
        if self.is_blender:
            ws += [self.timenet[0].weight, self.timenet[0].bias, self.timenet[2].weight, self.timenet[2].bias]
        for l in self.linear:
            ws += [l.weight, l.bias]
        out, hidden = _FusedMLP.apply(x, t, self.is_blender, int(hw.shape[0]), *ws, hw, hb)
        scaling = out[:, 3:5]
        if self.max_d_scale > 0:
            scaling = torch.tanh(scaling) * math.log(self.max_d_scale)
        ret = {'d_xyz': out[:, 0:3], 'd_rotation': out[:, 5:9], 'd_scaling': scaling, 'hidden': hidden, 'd_opacity': None, 'd_color': None}
        c = 9
        if self.local_frame:
            ret['local_rotation'] = out[:, c:c + 4]; c += 4
        if self.max_d_scale <= 0:
            # the un-sliced head matrix + column map, so ControlNodeWarp can blend without slicing copies
            self._packed = (out, (0, 5, 3, 9 if self.local_frame else -1))
        if self.pred_opacity:
            ret['d_opacity'] = out[:, c:c + 1]
        return ret

    def forward(self, x, t, **kwargs):
        if self._fusable(x, t):
            return self._forward_fused(x, t)
        return super().forward(x, t, **kwargs)


class _EagerDeformNetwork(nn.Module):
    """Reference: utils/time_utils.py:310-458 — constructor (parameter names / shapes / initialisers are the state-dict
    contract) and the eager layer sequence."""

    def __init__(self, D=8, W=256, input_ch=3, output_ch=59, t_multires=6, multires=10, is_blender=False,
                 local_frame=False, pred_opacity=False, pred_color=False, resnet_color=True, hash_color=False,
                 color_wrt_dir=False, progressive_brand_time=False, max_d_scale=-1, **kwargs):
        super().__init__()
        if pred_color or hash_color or progressive_brand_time:
            raise NotImplementedError("pred_color / hash_color / progressive_brand_time are outside the hot path (use the "
                                      "reference classes through d2gs_b200.install_into_reference())")
        self.name = 'mlp'
        self.D, self.W = D, W
        self.t_multires = 6 if is_blender else 10
        self.skips = [D // 2]
        self.progressive_brand_time = False
        self.embed_time_fn, time_input_ch = get_embedder(self.t_multires, 1)
        self.embed_fn, xyz_input_ch = get_embedder(multires, 3)
        self.input_ch = xyz_input_ch + time_input_ch
        self.pred_opacity, self.pred_color = pred_opacity, pred_color
        self.max_d_scale = max_d_scale
        self.reg_loss = 0.
        self.is_blender = is_blender
        if is_blender:
            self.time_out = 30
            self.timenet = nn.Sequential(nn.Linear(time_input_ch, 256), nn.ReLU(inplace=True), nn.Linear(256, self.time_out))
            in0 = xyz_input_ch + self.time_out
        else:
            in0 = self.input_ch
        self.linear = nn.ModuleList([nn.Linear(in0, W)] + [
            nn.Linear(W, W) if i not in self.skips else nn.Linear(W + in0, W) for i in range(D - 1)])
        self.gaussian_warp = nn.Linear(W, 3)
        self.gaussian_scaling = nn.Linear(W, 2)
        self.gaussian_rotation = nn.Linear(W, 4)
        self.local_frame = local_frame
        if self.local_frame:
            self.local_rotation = nn.Linear(W, 4)
            nn.init.normal_(self.local_rotation.weight, mean=0, std=1e-4)
            nn.init.zeros_(self.local_rotation.bias)
        for layer in self.linear:
            nn.init.kaiming_uniform_(layer.weight, mode='fan_in', nonlinearity='relu')
            nn.init.zeros_(layer.bias)
        nn.init.normal_(self.gaussian_warp.weight, mean=0, std=1e-5)
        nn.init.normal_(self.gaussian_scaling.weight, mean=0, std=1e-8)
        nn.init.normal_(self.gaussian_rotation.weight, mean=0, std=1e-5)
        nn.init.zeros_(self.gaussian_warp.bias)
        nn.init.zeros_(self.gaussian_scaling.bias)
        nn.init.zeros_(self.gaussian_rotation.bias)
        if self.pred_opacity:
            self.gaussian_opacity = nn.Linear(W, 1)
            nn.init.normal_(self.gaussian_opacity.weight, mean=0, std=1e-5)
            nn.init.zeros_(self.gaussian_opacity.bias)

    def trainable_parameters(self):
        return [{'params': list(self.parameters()), 'name': 'mlp'}]

    def forward(self, x, t, **kwargs):
        t_emb = self.embed_time_fn(t)
        if self.is_blender:
            t_emb = self.timenet(t_emb)
        x_emb = self.embed_fn(x)
        h = torch.cat([x_emb, t_emb], dim=-1)
        for i, l in enumerate(self.linear):
            h = F.relu(l(h))
            if i in self.skips:
                h = torch.cat([x_emb, t_emb, h], -1)
        scaling = self.gaussian_scaling(h)
        if self.max_d_scale > 0:
            scaling = torch.tanh(scaling) * math.log(self.max_d_scale)
        out = {'d_xyz': self.gaussian_warp(h), 'd_rotation': self.gaussian_rotation(h), 'd_scaling': scaling, 'hidden': h,
               'd_opacity': self.gaussian_opacity(h) if self.pred_opacity else None, 'd_color': None}
        if self.local_frame:
            out['local_rotation'] = self.local_rotation(h)
        return out

    def update(self, iteration, *args, **kwargs):
        return


class DeformNetwork(_FusedNetworkMixin, _EagerDeformNetwork):
    """Reference: utils/time_utils.py:310-458.  The dense layers run in the fused CUDA MLP on the M (<= a few thousand,
    or M*T for the regularisers) node rows; unsupported configurations fall through to the eager layers."""


class StaticNetwork(nn.Module):
    """Reference: utils/time_utils.py:288-307 (deform_type='static')."""

    def __init__(self, return_tensors=False, *args, **kwargs):
        super().__init__()
        self.name = 'static'
        self.reg_loss = 0.
        self.param = nn.Parameter(torch.zeros([1]))
        self.return_tensors = return_tensors

    def forward(self, x, *args, **kwargs):
        if self.return_tensors:
            z = lambda c: torch.zeros(x.shape[0], c, dtype=torch.float32, device=x.device)
            return {'d_xyz': z(3), 'd_rotation': z(4), 'd_scaling': z(2), 'hidden': z(1), 'd_opacity': None,
                    'd_color': None, 'local_rotation': z(4)}
        return {'d_xyz': 0., 'd_rotation': 0., 'd_scaling': 0., 'hidden': 0., 'd_opacity': None, 'd_color': None}

    def trainable_parameters(self):
        return [{'params': [self.param], 'name': 'deform'}]

    def update(self, *args, **kwargs):
        return


def farthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """(B,N,3) -> (B,npoint) indices; deterministic start at index 0 (init-time only, not on the hot path)."""
    B, N, _ = xyz.shape
    idx = torch.zeros((B, npoint), dtype=torch.long, device=xyz.device)
    dist = torch.full((B, N), 1e10, device=xyz.device)
    far = torch.zeros((B,), dtype=torch.long, device=xyz.device)
    b = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        idx[:, i] = far
        c = xyz[b, far][:, None, :]
        dist = torch.minimum(dist, ((xyz - c) ** 2).sum(-1))
        far = dist.argmax(-1)
    return idx


def landmark_interpolate(landmarks, steps, step, interpolation='log'):
    """Piecewise schedule of the regulariser weights (utils/time_utils.py:485-503): 0 before the first step, the last
    landmark after the last step, log- or linearly interpolated in between; a non-positive right landmark gives 0."""
    stage = sum(1 for s_ in steps if step >= s_)
    if stage == len(steps):
        return max(0, landmarks[-1])
    if stage == 0:
        return 0
    a, b = landmarks[stage - 1], landmarks[stage]
    if b <= 0:
        return 0
    r = (step - steps[stage - 1]) / (steps[stage] - steps[stage - 1])
    if interpolation == 'log':
        return math.exp(math.log(a) * (1 - r) + math.log(b) * r)
    if interpolation == 'linear':
        return a * (1 - r) + b * r
    raise NotImplementedError(f'Unknown interpolation type: {interpolation}')


def knn_torch(x: torch.Tensor, y: torch.Tensor, K: int):
    """K nearest rows of y for every row of x: (squared distances ascending (N,K), indices (N,K) int64).  Plain torch,
    exhaustive; ties resolved towards the lower index (the published semantics of pytorch3d.ops.knn_points, which the
    reference calls and this image does not have).  Used off the hot path only (regularisers, editing helpers)."""
    d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    K = min(K, y.shape[0])
    # stable sort = lower index first among equal distances
    order = torch.sort(d, dim=1, stable=True)
    return order.values[:, :K], order.indices[:, :K]


def arap_error_torch(nodes_seq: torch.Tensor, K: int = 10, radius: float = 0.1, least_edge_num: int = 3, sample_num: int = 512):
    """As-rigid-as-possible energy of a node trajectory (T, M, 3) against its first frame: the regulariser of
    ``ControlNodeWarp.arap_loss`` (utils/time_utils.py:1080-1089) = connectivity from the K nearest nodes of frame 0
    (``cal_connectivity_from_points``, utils/deform_utils.py:58-113: the first ``least_edge_num`` neighbours always, the
    others only inside ``radius``) + per-node best-fit rotation by SVD without gradient and the squared stretch of every
    edge (``cal_arap_error`` / ``estimate_rotation``, utils/deform_utils.py:130-205; unit edge weights, at most
    ``sample_num`` randomly drawn nodes).  Dense (M, K) formulation instead of the reference's edge lists."""
    T, M, _ = nodes_seq.shape
    p0 = nodes_seq[0]
    with torch.no_grad():
        d, idx = knn_torch(p0.detach(), p0.detach(), K + 1)
        d, idx = d[:, 1:], idx[:, 1:]                       # drop the node itself
        valid = torch.ones_like(d, dtype=torch.bool)
        valid[:, least_edge_num:] = d[:, least_edge_num:] < radius ** 2
        if M > sample_num:
            sel = torch.randint(0, M, (sample_num,), device=p0.device)
        else:
            sel = torch.arange(M, device=p0.device)
    w = valid[sel].to(p0.dtype)                              # (S, K) unit weights on existing edges
    edges = lambda v: (v[sel][:, None, :] - v[idx[sel]]) * w[..., None]      # (S, K, 3), zero where there is no edge
    src = edges(p0)
    err = p0.new_zeros(())
    for t in range(1, T):
        tgt = edges(nodes_seq[t])
        with torch.no_grad():
            S = torch.bmm(src.detach().transpose(1, 2), tgt.detach() * w[..., None])          # (S, 3, 3) covariance
            same = (src.detach() == tgt.detach()).all(dim=2).all(dim=1)
            S[same] = 0                                       # undeformed neighbourhood -> identity rotation
            U, sig, V = torch.svd(S)
            R = torch.bmm(V, U.transpose(1, 2))
            flip = torch.det(R) <= 0
            if bool(flip.any()):
                Um = U.clone()
                col = torch.argmin(sig[flip], dim=1)
                rows = torch.nonzero(flip, as_tuple=False).flatten()
                Um[rows, :, col] *= -1
                R[flip] = torch.bmm(V[flip], Um[flip].transpose(1, 2))
        rigid = torch.bmm(R, src.transpose(1, 2)).transpose(1, 2)
        err = err + (w * ((tgt - rigid) ** 2).sum(dim=2)).sum()
    return err


class _FastNodeWarpMixin:
    """B200 fast path of ``ControlNodeWarp``: ``forward`` and ``cal_nn_weight`` through the fused CUDA kernels
    (libd2gs.so).  Calls it does not cover — editing biases, skinning / hash-grid variants, external node sets — go to
    the next class in the MRO: the reference's own ControlNodeWarp when bound to it (``bind_reference``), else the
    stand-alone base below."""

    ORDER_REFRESH = 64   # forward calls between two rebuilds of the Morton processing order

    def _processing_order(self, x):
        """Cached spatial processing order of the surfels (speed only: a stale order is still a valid permutation, the
        centres move by a learning-rate step per iteration).  Rebuilt when the surfel count or device changes
        (densification / pruning) and every ORDER_REFRESH calls."""
        if not x.is_cuda:
            return None
        c = getattr(self, "_order_cache", None)
        P = int(x.shape[0])
        if c is None or c[0] != P or c[1] != x.device or c[3] >= self.ORDER_REFRESH:
            c = [P, x.device, processing_order(x), 0]
            object.__setattr__(self, "_order_cache", c)
        c[3] += 1
        return c[2]

    def _fast_variant(self) -> bool:
        return (self.d_rot_as_res and not getattr(self, "skinning", False) and not getattr(self, "use_hash", False)
                and not getattr(self, "pred_color", False) and not getattr(self, "cached_nn_weight", False))

    def cal_nn_weight(self, x, K=None, feature=None, nodes=None, gs_kernel=True, temperature=1.):
        """Reference: utils/time_utils.py:934-967.  Returns (nn_weight (P,K), nn_dist (P,K), nn_idx (P,K) int64)."""
        if not (gs_kernel and nodes is None and torch.is_tensor(x) and x.is_cuda and self._fast_variant()):
            return super().cal_nn_weight(x, K=K, feature=feature, nodes=nodes, gs_kernel=gs_kernel, temperature=temperature)
        K = self.K if K is None else K
        M = self.nodes.shape[0]
        zero = lambda c: torch.zeros((M, c), dtype=torch.float32, device=x.device)
        out = node_blend(x, feature, self.nodes, self._node_radius, self._node_weight.reshape(-1) if self.with_node_weight else None,
                         zero(3), zero(4), zero(2), None, None, K, self.hyper_dim)
        return out["nn_weight"], out["nn_dist"], out["nn_idx"]

    def forward(self, x, t, feature, motion_mask, iteration=0, is_training=True, node_trans_bias=None,
                node_scaling_bias=None, animation_d_values=None, **kwargs):
        if (node_trans_bias is not None or animation_d_values is not None or not torch.is_tensor(x) or not x.is_cuda
                or not self._fast_variant()):
            return super().forward(x, t, feature, motion_mask, iteration=iteration, is_training=is_training,
                                   node_trans_bias=node_trans_bias, node_scaling_bias=node_scaling_bias,
                                   animation_d_values=animation_d_values, **kwargs)
        if t.dim() == 0:
            t = self.expand_time(t)
        x = x.detach()
        net = self.network
        if hasattr(net, '_packed'):
            net._packed = None
        node_attrs = self.node_deform(t=t, **kwargs)
        packed = getattr(net, '_packed', None)
        wl = self._node_weight.reshape(-1) if self.with_node_weight else None
        mm = motion_mask if torch.is_tensor(motion_mask) else None
        order = self._processing_order(x)
        if packed is not None and t.dim() == 2 and packed[0].shape[0] == self.nodes.shape[0]:
            net._packed = None
            out = node_blend_packed(x, feature, self.nodes, self._node_radius, wl, packed[0], packed[1], mm, self.K, self.hyper_dim,
                                    order=order)
        else:
            out = node_blend(x, feature, self.nodes, self._node_radius, wl,
                             node_attrs['d_xyz'], node_attrs['d_rotation'], node_attrs['d_scaling'],
                             node_attrs.get('local_rotation') if self.local_frame else None, mm, self.K, self.hyper_dim,
                             order=order)
        ret = {'d_xyz': out['d_xyz'], 'd_rotation': out['d_rotation'], 'd_scaling': out['d_scaling'],
               'd_opacity': None, 'd_color': None}
        if self.pred_opacity:
            w, idx = out['nn_weight'], out['nn_idx']
            ret['d_opacity'] = (node_attrs['d_opacity'][idx] * w[..., None]).sum(dim=1) * motion_mask
        # regulariser schedule of the reference (utils/time_utils.py:1228-1232)
        self.reg_loss = 0.
        lam = landmark_interpolate(self.lambda_arap_landmarks, self.lambda_arap_steps, iteration)
        if self.training and lam > 0 and is_training:
            self.reg_loss = self.reg_loss + self.arap_loss() * lam
        return ret


class _StandaloneNodeWarp(nn.Module):
    """Reference: utils/time_utils.py:770-1233 without the reference tree: constructor / bookkeeping (parameter names and
    shapes are the state-dict contract), plain-torch versions of the calls outside the fast path.  The warm-up phase's
    node Gaussians (``as_gaussians``), node densification and the interactive-editing branches need the reference's
    ``scene`` package: bind to it with ``d2gs_b200.install_into_reference()`` and those members are the reference's own."""

    def __init__(self, is_blender, init_pcl=None, node_num=512, K=3, use_hash=False, hash_time=False,
                 enable_densify_prune=False, pred_opacity=False, pred_color=False, with_arap_loss=False,
                 with_node_weight=True, local_frame=False, d_rot_as_res=True, skinning=False, hyper_dim=2,
                 progressive_brand_time=False, max_d_scale=-1, is_scene_static=False, **kwargs):
        super().__init__()
        if use_hash or skinning or pred_color:
            raise NotImplementedError("use_hash / skinning / pred_color need the reference classes: d2gs_b200.install_into_reference()")
        self.K = K
        self.use_hash, self.hash_time = use_hash, hash_time
        self.enable_dp = enable_densify_prune
        self.name = 'node'
        self.with_node_weight = with_node_weight
        self.reg_loss = 0.
        self.local_frame = local_frame
        self.d_rot_as_res = d_rot_as_res
        self.hyper_dim = hyper_dim
        self.is_blender = is_blender
        self.pred_opacity, self.pred_color = pred_opacity, pred_color
        self.max_d_scale = max_d_scale
        self.is_scene_static = is_scene_static
        self.skinning = skinning
        self.with_arap_loss = with_arap_loss and not is_scene_static
        if self.with_arap_loss:      # utils/time_utils.py:790-795
            self.lambda_arap_landmarks = [1e-4, 1e-4, 1e-5, 1e-5, 0]
            self.lambda_arap_steps = [0, 5000, 10000, 20000, 20001]
        else:
            self.lambda_arap_landmarks, self.lambda_arap_steps = [0], [0]
        if self.is_scene_static:
            self.network = StaticNetwork(return_tensors=True)
        else:
            self.network = DeformNetwork(is_blender=is_blender, local_frame=local_frame, pred_opacity=pred_opacity,
                                         pred_color=pred_color, max_d_scale=max_d_scale)
        self.register_buffer('inited', torch.tensor(False))
        self.nodes = nn.Parameter(torch.randn(node_num, 3 + self.hyper_dim))
        self._node_radius = nn.Parameter(torch.randn(node_num))
        if self.with_node_weight:
            self._node_weight = nn.Parameter(torch.zeros_like(self.nodes[:, :1]), requires_grad=with_node_weight)
        self.cached_nn_weight = False
        self.nn_weight, self.nn_dist, self.nn_idxs = None, None, None
        self.gs = None

    # ---- bookkeeping identical to the reference ----
    def update(self, iteration):
        self.network.update(iteration)

    def trainable_parameters(self):
        node_params = [self.nodes, self._node_radius] + ([self._node_weight] if self.with_node_weight else [])
        return [{'params': list(self.network.parameters()), 'name': 'deform'}, {'params': node_params, 'name': 'nodes'}]

    @property
    def param_names(self):
        return ['nodes', '_node_radius', '_node_weight'] if self.with_node_weight else ['nodes', '_node_radius']

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        """utils/time_utils.py:845-866: node tables may change shape (densified checkpoints); ``gs_*`` entries are the
        node Gaussians of the warm-up phase — kept (``gs_state``) and re-emitted by ``state_dict`` so a checkpoint
        round-trips, applied to ``self.gs`` when one exists."""
        state_dict = dict(state_dict)
        for key in self.param_names:
            if key in state_dict:
                v = state_dict.pop(key)
                if getattr(self, key).shape != v.shape:
                    setattr(self, key, nn.Parameter(v))
                else:
                    getattr(self, key).data = v
        for key in [k for k in state_dict if k.startswith('gs_')]:
            v = state_dict.pop(key)
            self.gs_state[key] = v
            if self.gs is not None:
                try:
                    getattr(self.gs, key[3:]).data = v
                except Exception:
                    setattr(self.gs, key[3:], v)
        return super().load_state_dict(state_dict=state_dict, strict=False)

    @property
    def gs_state(self):
        d = self.__dict__.get('_gs_state')
        if d is None:
            d = {}
            object.__setattr__(self, '_gs_state', d)
        return d

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        if self.gs is not None and hasattr(self.gs, 'param_names'):
            for name in self.gs.param_names():
                sd['gs_' + name] = getattr(self.gs, name)
        else:
            for k, v in self.gs_state.items():
                sd[k] = v
        return sd

    @property
    def node_radius(self):
        return torch.exp(self._node_radius)

    @property
    def node_weight(self):
        return torch.sigmoid(self._node_weight)

    @property
    def node_num(self):
        return self.nodes.shape[0]

    def init(self, opt=None, init_pcl=None, hyper_pcl=None, keep_all=False, force_init=False, **kwargs):
        """Reference: utils/time_utils.py:884-927 (node Gaussians for the warm-up phase are not built here)."""
        if self.inited and not force_init:
            return
        dev = init_pcl.device
        self.inited.data = torch.ones_like(self.inited)
        if keep_all or self.node_num > init_pcl.shape[0]:
            self.nodes = nn.Parameter(torch.cat([init_pcl.float(), 1e-2 * torch.ones([init_pcl.shape[0], self.hyper_dim], device=dev)], dim=-1))
            init_nodes_idx = None
        else:
            pcl = init_pcl if hyper_pcl is None else hyper_pcl
            init_nodes_idx = farthest_point_sample(pcl.detach()[None], self.node_num)[0]
            self.nodes.data = torch.cat([init_pcl[init_nodes_idx].float(), 1e-2 * torch.ones([self.node_num, self.hyper_dim], device=dev)], dim=-1)
        scene_range = init_pcl.max() - init_pcl.min()
        r = torch.log(.1 * scene_range + 1e-7) * torch.ones([self.node_num], device=dev)
        if self._node_radius.shape != r.shape:
            self._node_radius = nn.Parameter(r)
            self._node_weight = nn.Parameter(torch.zeros_like(self.nodes[:, :1]))
        else:
            self._node_radius.data = r
            if self.with_node_weight:
                self._node_weight.data = torch.zeros_like(self.nodes[:, :1])
        return init_nodes_idx

    def expand_time(self, t):
        return t.unsqueeze(0).expand(self.nodes.shape[0], -1)

    def query_network(self, x, t, **kwargs):
        return self.network(x=x, t=t, **kwargs)

    def node_deform(self, t, detach_node=True, **kwargs):
        tshape = t.shape
        if t.dim() == 3:
            assert t.shape[0] == self.node_num
            nodes = self.nodes[:, None, ..., :3].expand(self.node_num, t.shape[1], 3).reshape(-1, 3)
            t = t.reshape(-1, 1)
        else:
            nodes = self.nodes[..., :3]
        if detach_node:
            nodes = nodes.detach()
        values = self.query_network(x=nodes, t=t, **kwargs)
        return {k: (v.view(*tshape[:-1], v.shape[-1]) if v is not None else None) for k, v in values.items()}

    def cal_nn_weight(self, x, K=None, feature=None, nodes=None, gs_kernel=True, temperature=1.):
        """Plain-torch version of utils/time_utils.py:934-967 for the calls the fused kernel does not take: CPU tensors,
        an external node set (``nodes=``), the softmax kernel (``gs_kernel=False``)."""
        if self.hyper_dim > 0 and feature is not None:
            x = torch.cat([x.detach(), feature[..., :self.hyper_dim]], dim=-1)
        K = self.K if K is None else K
        nd = self.nodes[..., :3].detach() if nodes is None else nodes[..., :3]
        if feature is not None:
            nd = torch.cat([nd[..., :3].detach(), self.nodes[..., 3:]], dim=-1)
        nn_dist, nn_idx = knn_torch(x, nd, K)
        if not gs_kernel:
            return torch.softmax(-nn_dist / temperature, dim=-1), nn_dist, nn_idx
        w = torch.exp(-nn_dist / (2 * self.node_radius[nn_idx] ** 2))
        if self.with_node_weight:
            w = w * self.node_weight[nn_idx][..., 0]
        w = w + 1e-7
        return w / w.sum(dim=-1, keepdim=True), nn_dist, nn_idx

    def forward(self, x, t, feature, motion_mask, iteration=0, is_training=True, node_trans_bias=None,
                node_scaling_bias=None, animation_d_values=None, **kwargs):
        raise NotImplementedError(
            "this call is outside the B200 fast path (CPU tensors, d_rot_as_res=False or the interactive-editing arguments "
            "node_trans_bias / animation_d_values): bind to the reference classes with d2gs_b200.install_into_reference()")

    def arap_loss(self, t=None, delta_t=0.05, t_samp_num=2):
        """utils/time_utils.py:1080-1089: ARAP energy of the nodes at ``t_samp_num`` random times in a window around t
        (an (M, T, 1) time batch through ``node_deform`` -> the fused MLP runs on M*T rows)."""
        dev = self.nodes.device
        t = torch.rand([], device=dev) if t is None else t.squeeze() + delta_t * (torch.rand([], device=dev) - .5)
        t_samp = torch.rand(t_samp_num, device=dev) * delta_t + t - .5 * delta_t
        t_samp = t_samp[None, :, None].expand(self.node_num, t_samp_num, 1)
        node_trans = self.node_deform(t=t_samp)['d_xyz']
        nodes_t = self.nodes[:, None, :3].detach() + node_trans      # (M, T, 3)
        return arap_error_torch(nodes_t.permute(1, 0, 2))

    def elastic_loss(self, t=None, delta_t=0.005, K=2, t_samp_num=8):
        """utils/time_utils.py:1091-1109."""
        dev = self.nodes.device
        t = torch.rand([], device=dev) if t is None else t.squeeze() + delta_t * (torch.rand([], device=dev) - .5)
        t_samp = torch.rand(t_samp_num, device=dev) * delta_t + t - .5 * delta_t
        t_samp = t_samp[None, :, None].expand(self.node_num, t_samp_num, 1)
        nodes_t = self.nodes[:, None, :3].detach() + self.node_deform(t=t_samp)['d_xyz']
        w, _, idx = self.cal_nn_weight(x=self.nodes[..., :3].detach(), feature=self.nodes[..., 3:], K=K + 1)
        w, idx = w[:, 1:], idx[:, 1:]
        var = (nodes_t[idx] - nodes_t[:, None]).norm(dim=-1).var(dim=2)
        var = var / (var.detach() + 1e-5)
        return (var * w).sum(dim=1).mean()

    def acc_loss(self, t=None, delta_t=.005):
        """utils/time_utils.py:1111-1122."""
        dev = self.nodes.device
        t = torch.rand([], device=dev) if t is None else t.squeeze() + delta_t * (torch.rand([], device=dev) - .5)
        t3 = torch.stack([t - delta_t, t, t + delta_t])[None, :, None].expand(self.node_num, 3, 1)
        nodes_t = self.nodes[:, None, :3].detach() + self.node_deform(t=t3)['d_xyz']
        acc = (nodes_t[:, 0] + nodes_t[:, 2] - 2 * nodes_t[:, 1]).norm(dim=-1)
        return (acc / (acc.detach() + 1e-5)).mean()

    @property
    def as_gaussians(self):
        raise NotImplementedError("node Gaussians of the warm-up phase are a scene.gaussian_model.StandardGaussianModel: "
                                  "bind to the reference tree with d2gs_b200.install_into_reference()")

    def densify(self, *a, **k):
        if not self.enable_dp and not k.get('force_dp', False):
            return                      # the reference returns silently when node densification is off (:1292-1293)
        raise NotImplementedError("node densification edits the reference trainer's optimiser state: bind to the reference "
                                  "classes with d2gs_b200.install_into_reference()")


class ControlNodeWarp(_FastNodeWarpMixin, _StandaloneNodeWarp):
    """Reference: utils/time_utils.py:770-1233.  Fast path = d_rot_as_res, no editing biases, no skinning/hash."""


model_dict = {'mlp': DeformNetwork, 'node': ControlNodeWarp, 'static': StaticNetwork}


class DeformModel:
    """Reference: scene/deform_model.py:13-72 (optimizer / LR schedule / checkpoint plumbing kept minimal)."""

    def __init__(self, deform_type='node', is_blender=False, d_rot_as_res=True, **kwargs):
        self.deform = model_dict[deform_type](is_blender=is_blender, d_rot_as_res=d_rot_as_res, **kwargs).cuda()
        self.name = self.deform.name
        self.optimizer = None
        self.spatial_lr_scale = 5
        self.d_rot_as_res = d_rot_as_res

    @property
    def reg_loss(self):
        return self.deform.reg_loss

    def step(self, xyz, time_emb, iteration=0, **kwargs):
        return self.deform(xyz, time_emb, iteration=iteration, **kwargs)

    def train_setting(self, training_args):
        lr = training_args.position_lr_init * self.spatial_lr_scale * training_args.deform_lr_scale
        groups = [{'params': g['params'], 'lr': lr, 'name': g['name']} for g in self.deform.trainable_parameters()]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)

    def save_weights(self, model_path, iteration):
        out = os.path.join(model_path, "deform/iteration_{}".format(iteration))
        os.makedirs(out, exist_ok=True)
        torch.save(self.deform.state_dict(), os.path.join(out, 'deform.pth'))

    def load_weights(self, model_path, iteration=-1):
        root = os.path.join(model_path, "deform")
        if iteration == -1:
            its = [int(f.split("_")[-1]) for f in os.listdir(root)] if os.path.isdir(root) else []
            if not its:
                return False
            iteration = max(its)
        path = os.path.join(root, "iteration_{}/deform.pth".format(iteration))
        if os.path.exists(path):
            self.deform.load_state_dict(torch.load(path))
            return True
        return False

    def densify(self, max_grad, x, x_grad, **kwargs):
        if self.name == 'node':
            self.deform.densify(max_grad=max_grad, optimizer=self.optimizer, x=x, x_grad=x_grad, **kwargs)

    def update(self, iteration):
        self.deform.update(iteration)


# --------------------------------------------------------------------------------------------------------------
# binding to the reference tree ("subclass of ControlNodeWarp overriding the forward fast path,
# falling back to the reference code otherwise; registered via scene.deform_model.model_dict['node']")
# --------------------------------------------------------------------------------------------------------------
_BOUND = {}


def bind_reference(time_utils_module):
    """Build the drop-in classes ON TOP of the reference's own (``utils.time_utils``): only ``forward`` /
    ``cal_nn_weight`` (ControlNodeWarp) and ``forward`` (DeformNetwork) are overridden, and only for the configurations
    the CUDA kernels cover; ``arap_loss``, ``densify``, ``as_gaussians``, ``init`` (incl. the node Gaussians), the
    editing branches, ``state_dict`` with its ``gs_*`` keys ... are inherited from the reference unchanged."""
    key = id(time_utils_module)
    if key in _BOUND:
        return _BOUND[key]
    ref = time_utils_module

    class RefDeformNetwork(_FusedNetworkMixin, ref.DeformNetwork):
        pass

    class RefControlNodeWarp(_FastNodeWarpMixin, ref.ControlNodeWarp):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            if type(self.network) is ref.DeformNetwork:      # same parameters, fused forward
                self.network.__class__ = RefDeformNetwork

    for c, n in ((RefDeformNetwork, "DeformNetwork"), (RefControlNodeWarp, "ControlNodeWarp")):
        c.__name__ = c.__qualname__ = n
    out = {'mlp': RefDeformNetwork, 'node': RefControlNodeWarp, 'static': ref.StaticNetwork}
    _BOUND[key] = out
    return out


def install_into_reference(patch_renderer: bool = True):
    """Make an importable reference tree use the B200 path without editing it:
      * ``scene.deform_model.model_dict`` (scene/deform_model.py:10) gets the classes of ``bind_reference``;
      * if the ``gaussian_renderer`` that is importable is the REFERENCE's package (this repo's shadow package is not
        first on sys.path), its ``render`` / ``render_flow`` are replaced by the B200 ones, so
        ``from gaussian_renderer import render`` (train_gui.py:18, render_mesh.py:16) picks them up.
    Returns the ``scene.deform_model`` module."""
    import importlib
    import importlib.util
    tu = importlib.import_module("utils.time_utils")
    ref_dm = importlib.import_module("scene.deform_model")
    ref_dm.model_dict.update(bind_reference(tu))
    if patch_renderer:
        gr = importlib.import_module("gaussian_renderer")
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if not os.path.abspath(getattr(gr, "__file__", "") or "").startswith(here):
            # the reference's package: load this repo's implementation under a private name and graft its entry points
            spec = importlib.util.spec_from_file_location("d2gs_b200._gaussian_renderer",
                                                          os.path.join(here, "gaussian_renderer", "__init__.py"))
            ours = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(ours)
            gr.render_reference, gr.render_flow_reference = gr.render, getattr(gr, "render_flow", None)
            gr.render, gr.render_flow = ours.render, ours.render_flow
    return ref_dm
