"""Node-controlled deformation, B200 path: the time-conditioned MLP evaluated at the M control nodes, then the
fused CUDA KNN + radial-basis weighting + blend onto the P surfels (libd2gs.so: d2gs_deform_forward/backward).

Mirrors the reference interface (paths relative to /root/reference):
  * ``DeformNetwork``      utils/time_utils.py:310-458   same constructor, parameter names and output dict
  * ``ControlNodeWarp``    utils/time_utils.py:770-1233  same constructor, ``forward(x, t, feature, motion_mask, ...)``,
                           ``cal_nn_weight``, ``node_deform``, ``query_network``, ``expand_time``, state-dict keys
                           (``nodes``, ``_node_radius``, ``_node_weight``, ``network.*``, buffer ``inited``)
  * ``DeformModel``        scene/deform_model.py:13-72   ``step(xyz, time_emb, iteration=0, **kwargs)``
``install_into_reference()`` registers the classes in ``scene.deform_model.model_dict`` so train_gui.py runs unchanged.

Out of scope here: ARAP regularisers, node densification/pruning, the interactive
editing branches (node_trans_bias / animation_d_values), hash-grid and skinning variants — they raise.
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


class DeformNetwork(nn.Module):
    """Reference: utils/time_utils.py:310-458.  The dense layers are plain GEMMs on M (<= a few thousand) rows."""

    def __init__(self, D=8, W=256, input_ch=3, output_ch=59, t_multires=6, multires=10, is_blender=False,
                 local_frame=False, pred_opacity=False, pred_color=False, resnet_color=True, hash_color=False,
                 color_wrt_dir=False, progressive_brand_time=False, max_d_scale=-1, **kwargs):
        super().__init__()
        if pred_color or hash_color or progressive_brand_time:
            raise NotImplementedError("pred_color / hash_color / progressive_brand_time are outside the hot path")
        self.name = 'mlp'
        self.D, self.W = D, W
        self.t_multires = 6 if is_blender else 10
        self.skips = [D // 2]
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

    use_fused = True   # set False to force the eager torch layers (tests compare the two)

    def _fusable(self, x, t) -> bool:
        return (self.use_fused and x.is_cuda and self.D == 8 and self.W == 256 and self.skips == [4] and x.dim() == 2 and x.shape[1] == 3
                and isinstance(self.embed_fn, Embedder) and self.embed_fn.multires == 10 and t.shape[-1] == 1
                and (t.numel() == x.shape[0] or t.numel() == 1))

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


class ControlNodeWarp(nn.Module):
    """Reference: utils/time_utils.py:770-1233.  Fast path = d_rot_as_res, no editing biases, no skinning/hash."""

    def __init__(self, is_blender, init_pcl=None, node_num=512, K=3, use_hash=False, hash_time=False,
                 enable_densify_prune=False, pred_opacity=False, pred_color=False, with_arap_loss=False,
                 with_node_weight=True, local_frame=False, d_rot_as_res=True, skinning=False, hyper_dim=2,
                 progressive_brand_time=False, max_d_scale=-1, is_scene_static=False, **kwargs):
        super().__init__()
        if use_hash or skinning or pred_color or not d_rot_as_res:
            raise NotImplementedError("use_hash / skinning / pred_color / d_rot_as_res=False are outside the B200 hot path")
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
        state_dict = dict(state_dict)
        for key in self.param_names:
            if key in state_dict:
                v = state_dict.pop(key)
                if getattr(self, key).shape != v.shape:
                    setattr(self, key, nn.Parameter(v))
                else:
                    getattr(self, key).data = v
        for key in [k for k in state_dict if k.startswith('gs_')]:
            state_dict.pop(key)   # node Gaussians of the warm-up phase are not part of the hot path
        return super().load_state_dict(state_dict=state_dict, strict=False)

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

    def cal_nn_weight(self, x, K=None, feature=None, nodes=None, gs_kernel=True, temperature=1.):
        """Reference: utils/time_utils.py:934-967.  Returns (nn_weight (P,K), nn_dist (P,K), nn_idx (P,K) int64)."""
        if not gs_kernel or nodes is not None:
            raise NotImplementedError("softmax kernel / external node sets belong to the editing branches")
        K = self.K if K is None else K
        M = self.nodes.shape[0]
        zero = lambda c: torch.zeros((M, c), dtype=torch.float32, device=x.device)
        out = node_blend(x, feature, self.nodes, self._node_radius, self._node_weight.reshape(-1) if self.with_node_weight else None,
                         zero(3), zero(4), zero(2), None, None, K, self.hyper_dim)
        return out["nn_weight"], out["nn_dist"], out["nn_idx"]

    def forward(self, x, t, feature, motion_mask, iteration=0, is_training=True, node_trans_bias=None,
                node_scaling_bias=None, animation_d_values=None, **kwargs):
        if node_trans_bias is not None or animation_d_values is not None:
            raise NotImplementedError("interactive-editing branches (node_trans_bias / animation_d_values) are out of scope")
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
        self.reg_loss = 0.
        return ret

    def arap_loss(self, *a, **k):
        raise NotImplementedError("ARAP regularisers (utils/deform_utils.py) are outside the hot path")

    def densify(self, *a, **k):
        raise NotImplementedError("node densification is outside the hot path")


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

    def update(self, iteration):
        self.deform.update(iteration)


def install_into_reference():
    """Register the B200 classes into the reference's ``scene.deform_model.model_dict`` (scene/deform_model.py:10)."""
    import scene.deform_model as ref_dm   # the reference package must be importable
    ref_dm.model_dict['node'] = ControlNodeWarp
    ref_dm.model_dict['mlp'] = DeformNetwork
    ref_dm.model_dict['static'] = StaticNetwork
    return ref_dm
