"""TEST INFRASTRUCTURE — torch-CPU restatement of the reference's node-controlled deformation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this.

Follows (paths relative to /root/reference):
  utils/time_utils.py:208-256   positional embedder (include_input, log-sampled 2^k frequencies, [sin, cos])
  utils/time_utils.py:310-453   DeformNetwork (timenet 13->256->30, 8x256 trunk with skip after layer 4, heads)
  utils/time_utils.py:934-967   cal_nn_weight
  utils/time_utils.py:115-132   quaternion_to_matrix (un-normalised, 2/|q|^2)
  utils/time_utils.py:1133-1233 ControlNodeWarp.forward (d_rot_as_res=True fast path)

PARITY UNPINNED for the K-NN: the reference calls ``pytorch3d.ops.knn_points`` (utils/time_utils.py:950);
pytorch3d is neither vendored nor version-pinned (requirements.txt: a local path; readme.md:61 installs git HEAD)
and is not installed here.  Its published semantics are restated: squared L2 distances, the K smallest in
ascending order, int64 indices, differentiable w.r.t. both point sets.  Ties: lower node index first.
Everything else in this file is plain torch arithmetic and differentiable through torch autograd, which is how the
gradients of the CUDA path are checked.

PIN for everything but the K-NN: tests/golden/deform_golden.npz holds outputs and gradients of the reference's OWN
``ControlNodeWarp`` class, imported from /root/reference/utils/time_utils.py and run on CPU by
tests/golden/make_deform_golden.py (pytorch3d stubbed as above, ``Module.cuda`` a no-op); ``init_network_params`` loads into
the reference's DeformNetwork with ``strict=True``, i.e. names and shapes are the reference's.  tests/test_deform_cpu.py
checks this restatement against it (outputs 2e-6, gradients 1e-5), tests/test_deform_gpu.py the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def embed(x: torch.Tensor, multires: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  (time_utils.py:208-256)."""
    outs = [x]
    freqs = (2.0 ** torch.linspace(0.0, multires - 1, steps=multires)).tolist()
    for f in freqs:
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def init_network_params(seed: int = 0, D: int = 8, W: int = 256, multires: int = 10, t_multires: int = 6,
                        local_frame: bool = True, head_scale: float = 1.0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference initialisation of DeformNetwork(is_blender=True) (time_utils.py:344-382), seeded.

    ``head_scale`` multiplies the (tiny) default head weights so that deltas are non-trivial in tests
    (SURVEY.md §8(d) "heads scaled x1e3" variant)."""
    g = torch.Generator().manual_seed(seed)
    xyz_ch, t_ch, t_out = 3 + 3 * 2 * multires, 1 + 2 * t_multires, 30
    p: Dict[str, torch.Tensor] = {}

    def default_linear(name, fan_in, fan_out):   # nn.Linear default init
        bound = 1 / math.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand((fan_out, fan_in), generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand((fan_out,), generator=g) * 2 - 1) * bound

    def kaiming(name, fan_in, fan_out):          # kaiming_uniform_(mode=fan_in, relu), zero bias
        bound = math.sqrt(6.0 / fan_in)
        p[name + ".weight"] = (torch.rand((fan_out, fan_in), generator=g) * 2 - 1) * bound
        p[name + ".bias"] = torch.zeros(fan_out)

    default_linear("timenet.0", t_ch, 256)
    default_linear("timenet.2", 256, t_out)
    in0 = xyz_ch + t_out
    skips = [D // 2]
    kaiming("linear.0", in0, W)
    for i in range(D - 1):
        kaiming(f"linear.{i + 1}", W + in0 if i in skips else W, W)

    def head(name, out, std):
        p[name + ".weight"] = torch.randn((out, W), generator=g) * std * head_scale
        p[name + ".bias"] = torch.zeros(out)

    head("gaussian_warp", 3, 1e-5)
    head("gaussian_scaling", 2, 1e-8)
    head("gaussian_rotation", 4, 1e-5)
    if local_frame:
        head("local_rotation", 4, 1e-4)
    return {k: v.to(dtype) for k, v in p.items()}


def deform_network_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, t: torch.Tensor, D: int = 8,
                           multires: int = 10, t_multires: int = 6) -> Dict[str, torch.Tensor]:
    """DeformNetwork.forward for is_blender=True (time_utils.py:410-453)."""
    t_emb = embed(t, t_multires)
    t_emb = F.linear(F.relu(F.linear(t_emb, p["timenet.0.weight"], p["timenet.0.bias"])), p["timenet.2.weight"],
                     p["timenet.2.bias"])
    x_emb = embed(x, multires)
    h = torch.cat([x_emb, t_emb], -1)
    skips = [D // 2]
    for i in range(D):
        h = F.relu(F.linear(h, p[f"linear.{i}.weight"], p[f"linear.{i}.bias"]))
        if i in skips:
            h = torch.cat([x_emb, t_emb, h], -1)
    out = {"d_xyz": F.linear(h, p["gaussian_warp.weight"], p["gaussian_warp.bias"]),
           "d_scaling": F.linear(h, p["gaussian_scaling.weight"], p["gaussian_scaling.bias"]),
           "d_rotation": F.linear(h, p["gaussian_rotation.weight"], p["gaussian_rotation.bias"]), "hidden": h}
    if "local_rotation.weight" in p:
        out["local_rotation"] = F.linear(h, p["local_rotation.weight"], p["local_rotation.bias"])
    return out


def knn_points(x: torch.Tensor, nodes: torch.Tensor, K: int, mode: str = "exact"):
    """Published semantics of pytorch3d.ops.knn_points: squared L2, K smallest ascending, int64 idx.

    mode="exact": explicit (x-n)^2 sums, stable sort (the parity oracle; materialises P*M*D, small cases only).
    mode="mm":    |x|^2+|n|^2-2x.n via one GEMM + topk — the memory-sane stand-in used to TIME the reference
                  pipeline at BASELINE sizes (pytorch3d's fused brute-force kernel is not available here)."""
    if mode == "mm":
        d2 = ((x * x).sum(1, keepdim=True) + (nodes * nodes).sum(1)[None, :] - 2.0 * (x @ nodes.T)).clamp_min(0)
        dist, idx = torch.topk(d2, K, dim=1, largest=False, sorted=True)
        return dist, idx
    if x.shape[0] * nodes.shape[0] * x.shape[1] > (1 << 28):
        # same selection at BASELINE sizes without materialising P*M*D at once: indices chunk by chunk (no gradient
        # flows through a selection), then the K distances recomputed with the same explicit arithmetic
        with torch.no_grad():
            order = torch.cat([torch.sort(((x[s:s + 8192, None, :] - nodes[None, :, :]) ** 2).sum(-1), dim=1, stable=True).indices[:, :K]
                               for s in range(0, x.shape[0], 8192)])
        return ((x[:, None, :] - nodes[order]) ** 2).sum(-1), order
    d2 = ((x[:, None, :] - nodes[None, :, :]) ** 2).sum(-1)   # explicit: cdist()**2 is not bit-identical
    order = torch.sort(d2, dim=1, stable=True).indices[:, :K]
    return torch.gather(d2, 1, order), order


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def cal_nn_weight(x, feature, nodes, node_radius_log, node_weight_logit, K, hyper_dim, knn_mode="exact"):
    """time_utils.py:934-967 (gs_kernel=True, not skinning, no cache)."""
    if hyper_dim > 0 and feature is not None:
        xq = torch.cat([x.detach(), feature[..., :hyper_dim]], -1)
        nq = torch.cat([nodes[..., :3].detach(), nodes[..., 3:]], -1)
    else:
        xq = x.detach()
        nq = nodes[..., :3].detach()
    nn_dist, nn_idx = knn_points(xq, nq, K, knn_mode)
    radius = torch.exp(node_radius_log)[nn_idx]
    w = torch.exp(-nn_dist / (2 * radius ** 2))
    if node_weight_logit is not None:
        w = w * torch.sigmoid(node_weight_logit)[nn_idx][..., 0]
    w = w + 1e-7
    w = w / w.sum(-1, keepdim=True)
    return w, nn_dist, nn_idx


def blend(x, nn_weight, nn_idx, nodes, node_attrs, motion_mask, local_frame):
    """time_utils.py:1145-1157, 1190-1194 (d_rot_as_res=True)."""
    x = x.detach()
    node_trans, node_rot, node_scale = node_attrs["d_xyz"], node_attrs["d_rotation"], node_attrs["d_scaling"]
    if local_frame:
        rot_bias = torch.tensor([1.0, 0, 0, 0], dtype=x.dtype, device=x.device)
        Rm = quaternion_to_matrix(node_attrs["local_rotation"] + rot_bias)
        nn_nodes = nodes[nn_idx][..., :3].detach()
        Ax = torch.einsum("nkab,nkb->nka", Rm[nn_idx], x[:, None] - nn_nodes) + nn_nodes + node_trans[nn_idx]
        translate = (Ax * nn_weight[..., None]).sum(1) - x
    else:
        translate = (node_trans[nn_idx] * nn_weight[..., None]).sum(1)
    translate = translate * motion_mask
    rotation = (node_rot[nn_idx] * nn_weight[..., None]).sum(1) * motion_mask
    scale = (node_scale[nn_idx] * nn_weight[..., None]).sum(1) * motion_mask
    return {"d_xyz": translate, "d_rotation": rotation, "d_scaling": scale}


def control_node_warp_forward(net_params, nodes, node_radius_log, node_weight_logit, x, t, feature, motion_mask, K,
                              hyper_dim, local_frame=True, D=8, knn_mode="exact"):
    """ControlNodeWarp.forward (time_utils.py:1133-1233), fast path.  ``t``: (M,1) time per node."""
    w, nn_dist, nn_idx = cal_nn_weight(x, feature, nodes, node_radius_log, node_weight_logit, K, hyper_dim, knn_mode)
    attrs = deform_network_forward(net_params, nodes[..., :3].detach(), t, D=D)
    out = blend(x, w, nn_idx, nodes, attrs, motion_mask, local_frame)
    out.update(nn_weight=w, nn_dist=nn_dist, nn_idx=nn_idx, node_attrs=attrs)
    return out


def render_glue_pre(xyz, scaling, rotation, opacity, d_xyz, d_rotation, d_scaling):
    """gaussian_renderer/__init__.py:83-99 with the activations of scene/gaussian_model.py:67-75,101-124."""
    means3D = xyz + d_xyz
    opac = torch.sigmoid(opacity)
    scales = torch.exp(scaling) + d_scaling
    rot = F.normalize(rotation + d_rotation)
    return means3D, opac, scales, rot
