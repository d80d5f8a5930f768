"""Generates tests/golden/deform_golden.npz with the REFERENCE's own node-controlled deformation code: the class
``ControlNodeWarp`` of /root/reference/utils/time_utils.py (DeformNetwork.forward :410-453, cal_nn_weight :934-967,
node_deform :990-1002, forward :1133-1233) is imported in the build container and run on CPU tensors — no GPU needed.
Committed together with its output; nothing at test time reads /root/reference.

What is NOT the reference here, and why: ``pytorch3d`` is not installed (and neither vendored nor version-pinned by the
reference), so ``pytorch3d.ops.knn_points`` is replaced by its published semantics (squared L2, K smallest ascending, lower
index on ties, int64 indices, differentiable) — the K-NN stays "parity unpinned"; everything downstream of it (radial-basis
weights, DeformNetwork, local-frame blend, all gradients) is the reference's code.  ``Module.cuda()`` is made a no-op so the
constructor runs on CPU.  The network weights are not the reference's random initialisation but the seeded tensors of
oracle/deform_oracle.init_network_params (heads x1e3 so that the deltas are not ~1e-5), loaded with
``load_state_dict(strict=True)`` — which also proves that the oracle's parameter names and shapes are the reference's.

Inputs are re-drawn by the tests from deform_case(); only outputs and gradients are stored."""
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

CASES = {"local": dict(P=3000, M=64, K=4, hyper=8, local_frame=True, seed=21, fid=0.37),
         "plain": dict(P=2000, M=96, K=3, hyper=8, local_frame=False, seed=22, fid=0.81)}


def deform_case(name):
    """Seeded inputs of one case (numpy / torch CPU): surfel centres, hyper features, nodes, radii, weights, upstream gradients."""
    from d2gs_b200 import synthetic as syn
    from oracle import deform_oracle as do
    c = CASES[name]
    sc = syn.make_scene(c["P"], c["seed"], 0.01, n_nodes=c["M"], hyper_dim=c["hyper"])
    g = torch.Generator().manual_seed(c["seed"])
    nodes = torch.as_tensor(sc.nodes).clone()
    nodes[:, 3:] += 0.02 * torch.randn(c["M"], c["hyper"], generator=g)          # distinct hyper coordinates
    feature = torch.as_tensor(sc.feature).clone() + 0.02 * torch.randn(c["P"], c["hyper"], generator=g)
    out = dict(c, xyz=torch.as_tensor(sc.xyz), feature=feature, nodes=nodes,
               node_radius=torch.as_tensor(sc.node_radius) + 0.1 * torch.randn(c["M"], generator=g),
               node_weight=0.5 * torch.randn(c["M"], 1, generator=g),
               net=do.init_network_params(seed=c["seed"], local_frame=c["local_frame"], head_scale=1e3),
               g_xyz=torch.randn(c["P"], 3, generator=g), g_rot=torch.randn(c["P"], 4, generator=g), g_scale=torch.randn(c["P"], 2, generator=g))
    return out


def _knn_points(p1, p2, lengths1=None, lengths2=None, K=1, **kw):
    """Published semantics of pytorch3d.ops.knn_points for one batch element (see the module docstring)."""
    d2 = ((p1[0][:, None, :] - p2[0][None, :, :]) ** 2).sum(-1)
    order = torch.sort(d2.detach(), dim=1, stable=True).indices[:, :K]
    return torch.gather(d2, 1, order)[None], order[None], None


def _import_reference():
    pt3 = types.ModuleType("pytorch3d")
    ops = types.ModuleType("pytorch3d.ops"); ops.knn_points = _knn_points; ops.ball_query = None
    loss = types.ModuleType("pytorch3d.loss"); mls = types.ModuleType("pytorch3d.loss.mesh_laplacian_smoothing"); mls.cot_laplacian = None
    io = types.ModuleType("pytorch3d.io"); io.load_ply = None
    pt3.ops, pt3.loss, pt3.io, loss.mesh_laplacian_smoothing = ops, loss, io, mls
    sys.modules.update({"pytorch3d": pt3, "pytorch3d.ops": ops, "pytorch3d.loss": loss, "pytorch3d.loss.mesh_laplacian_smoothing": mls,
                        "pytorch3d.io": io})
    sys.path.insert(0, "/root/reference")
    import utils.time_utils as tu      # noqa: E402  (reference code, read-only)
    return tu


if __name__ == "__main__":
    tu = _import_reference()
    torch.nn.Module.cuda = lambda self, *a, **k: self
    out = {}
    for name in CASES:
        c = deform_case(name)
        model = tu.ControlNodeWarp(is_blender=True, node_num=c["M"], K=c["K"], with_node_weight=True, local_frame=c["local_frame"],
                                   hyper_dim=c["hyper"], d_rot_as_res=True)
        model.network.load_state_dict(c["net"], strict=True)
        with torch.no_grad():
            model.nodes.copy_(c["nodes"]); model._node_radius.copy_(c["node_radius"]); model._node_weight.copy_(c["node_weight"])
        feature = c["feature"].clone().requires_grad_(True)
        t = torch.full((c["M"], 1), c["fid"])
        d = model(c["xyz"], t, feature, torch.ones(c["P"], 1))
        loss = (d["d_xyz"] * c["g_xyz"]).sum() + (d["d_rotation"] * c["g_rot"]).sum() + (d["d_scaling"] * c["g_scale"]).sum()
        loss.backward()
        for k in ("d_xyz", "d_rotation", "d_scaling"):
            out[f"{name}_{k}"] = d[k].detach().numpy()
        out[f"{name}_g_feature"] = feature.grad.numpy()
        out[f"{name}_g_nodes"] = model.nodes.grad.numpy()
        out[f"{name}_g_node_radius"] = model._node_radius.grad.numpy()
        out[f"{name}_g_node_weight"] = model._node_weight.grad.numpy()
        for k, p in model.network.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            if p.numel() <= 4096 or k.startswith("gaussian_") or k.startswith("local_"):
                out[f"{name}_g_net_{k}"] = g.numpy()
            out[f"{name}_gnorm_net_{k}"] = np.float64(g.double().norm().item())
        print(name, {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith(name) and not k.startswith(name + "_g")})
    np.savez_compressed(os.path.join(HERE, "deform_golden.npz"), **out)
    print("wrote deform_golden.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")
