"""CPU: the deformation oracle's structure (it is the yardstick of the CUDA deform path) and the host-side mirror
of the reference interface (parameter names, shapes, state-dict round trip) — no CUDA needed."""
import math

import numpy as np
import pytest
import torch

import util  # noqa: F401
from oracle import deform_oracle as do


def test_embedder_and_network_shapes():
    x = torch.randn(7, 3)
    assert do.embed(x, 10).shape == (7, 63) and do.embed(torch.rand(7, 1), 6).shape == (7, 13)
    e = do.embed(x, 2)
    assert torch.equal(e[:, :3], x) and torch.allclose(e[:, 3:6], torch.sin(x)) and torch.allclose(e[:, 9:12], torch.sin(2 * x))
    p = do.init_network_params(seed=0, local_frame=True)
    n_params = sum(v.numel() for k, v in p.items())
    assert n_params == 523051                      # SURVEY.md §8 a17 (with the local_rotation head)
    assert p["linear.5.weight"].shape == (256, 349)   # skip-concat after layer 4
    out = do.deform_network_forward(p, x, torch.rand(7, 1))
    assert out["d_xyz"].shape == (7, 3) and out["d_rotation"].shape == (7, 4) and out["d_scaling"].shape == (7, 2)
    assert out["local_rotation"].shape == (7, 4)


def test_knn_modes_agree_and_are_sorted():
    g = torch.Generator().manual_seed(0)
    x, n = torch.randn(500, 11, generator=g), torch.randn(64, 11, generator=g)
    d, i = do.knn_points(x, n, 4)
    d2, i2 = do.knn_points(x, n, 4, mode="mm")
    assert (d[:, 1:] >= d[:, :-1]).all() and i.dtype == torch.int64
    assert (i == i2).float().mean() > 0.99 and torch.allclose(d, d2, atol=1e-4)
    brute = ((x[:, None] - n[None]) ** 2).sum(-1)
    assert torch.equal(torch.gather(brute, 1, i), d)


def test_weights_normalised_and_quaternion_matrix():
    g = torch.Generator().manual_seed(1)
    x, feat = torch.randn(300, 3, generator=g), torch.randn(300, 8, generator=g) * 0.01
    nodes = torch.cat([torch.randn(32, 3, generator=g), torch.full((32, 8), 1e-2)], 1)
    w, d, i = do.cal_nn_weight(x, feat, nodes, torch.full((32,), math.log(0.3)), torch.zeros(32, 1), 4, 8)
    assert torch.allclose(w.sum(-1), torch.ones(300), atol=1e-6) and (w > 0).all()
    q = torch.randn(10, 4, generator=g)
    R = do.quaternion_to_matrix(q)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(10, 3, 3), atol=1e-5)   # orthonormal for any |q|
    assert torch.allclose(do.quaternion_to_matrix(torch.tensor([[2.0, 0, 0, 0]])), torch.eye(3)[None])


def test_identity_deformation_is_zero():
    """zero heads => zero translation in the local frame too (R = I, Ax = x + 0)."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(100, 3, generator=g)
    nodes = torch.cat([torch.randn(16, 3, generator=g), torch.full((16, 8), 1e-2)], 1)
    attrs = {"d_xyz": torch.zeros(16, 3), "d_rotation": torch.zeros(16, 4), "d_scaling": torch.zeros(16, 2),
             "local_rotation": torch.zeros(16, 4)}
    w, _, i = do.cal_nn_weight(x, None, nodes, torch.zeros(16), None, 3, 8)
    out = do.blend(x, w, i, nodes, attrs, torch.ones(100, 1), True)
    assert out["d_xyz"].abs().max() < 1e-6 and not out["d_rotation"].any()


def test_host_side_interface_mirror():
    from d2gs_b200 import deform as dfm
    net = dfm.DeformNetwork(is_blender=True, local_frame=True)
    names = set(dict(net.named_parameters()))
    ref = do.init_network_params(local_frame=True)
    assert names == set(ref)                                   # same parameter names as the reference state dict
    for k, v in net.named_parameters():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    out = net(torch.randn(5, 3), torch.rand(5, 1))
    assert set(out) >= {"d_xyz", "d_rotation", "d_scaling", "hidden", "d_opacity", "d_color", "local_rotation"}
    # numerics of the module == the oracle with the same weights
    o2 = do.deform_network_forward({k: v.detach() for k, v in net.named_parameters()}, torch.ones(5, 3) * 0.3, torch.full((5, 1), 0.25))
    o1 = net(torch.ones(5, 3) * 0.3, torch.full((5, 1), 0.25))
    for k in ("d_xyz", "d_rotation", "d_scaling", "local_rotation"):
        assert torch.allclose(o1[k], o2[k], atol=1e-7), k
    cn = dfm.ControlNodeWarp(is_blender=True, node_num=32, K=4, hyper_dim=8, local_frame=True)
    sd = cn.state_dict()
    assert {"nodes", "_node_radius", "_node_weight", "inited"} <= set(sd) and any(k.startswith("network.linear.") for k in sd)
    cn2 = dfm.ControlNodeWarp(is_blender=True, node_num=16, K=4, hyper_dim=8, local_frame=True)
    cn2.load_state_dict(sd)                                    # node count follows the checkpoint, like the reference
    assert cn2.nodes.shape == (32, 11) and cn2.node_num == 32
    assert [g["name"] for g in cn.trainable_parameters()] == ["deform", "nodes"]
    assert cn.expand_time(torch.tensor([0.5])).shape == (32, 1)
    with pytest.raises(NotImplementedError):
        dfm.ControlNodeWarp(is_blender=True, use_hash=True)
    assert set(dfm.model_dict) == {"mlp", "node", "static"}
    idx = dfm.farthest_point_sample(torch.randn(1, 200, 3), 10)
    assert idx.shape == (1, 10) and len(set(idx[0].tolist())) == 10


@pytest.mark.parametrize("name", ["local", "plain"])
def test_oracle_matches_the_references_own_control_node_warp(name):
    """tests/golden/deform_golden.npz was produced by the reference's own ControlNodeWarp class (utils/time_utils.py, run on
    CPU by tests/golden/make_deform_golden.py with pytorch3d.ops.knn_points replaced by its published semantics): outputs and
    every gradient of the oracle's restatement must agree with it."""
    import os
    import sys
    sys.path.insert(0, os.path.join(util.ROOT, "tests", "golden"))
    from make_deform_golden import deform_case
    g = np.load(os.path.join(util.ROOT, "tests", "golden", "deform_golden.npz"))
    c = deform_case(name)
    net = {k: v.clone().requires_grad_(True) for k, v in c["net"].items()}
    nodes, rad, wl = (c[k].clone().requires_grad_(True) for k in ("nodes", "node_radius", "node_weight"))
    feature = c["feature"].clone().requires_grad_(True)
    t = torch.full((c["M"], 1), c["fid"])
    out = do.control_node_warp_forward(net, nodes, rad, wl, c["xyz"], t, feature, torch.ones(c["P"], 1), c["K"], c["hyper"],
                                       local_frame=c["local_frame"])
    for k in ("d_xyz", "d_rotation", "d_scaling"):
        assert util.rel_err(out[k].detach().numpy(), g[f"{name}_{k}"]) < 2e-6, k
    ((out["d_xyz"] * c["g_xyz"]).sum() + (out["d_rotation"] * c["g_rot"]).sum() + (out["d_scaling"] * c["g_scale"]).sum()).backward()
    for k, v in (("g_feature", feature), ("g_nodes", nodes), ("g_node_radius", rad), ("g_node_weight", wl)):
        assert util.rel_err(v.grad.numpy(), g[f"{name}_{k}"]) < 1e-5, k
    for k, p in net.items():
        gn = float(p.grad.double().norm()) if p.grad is not None else 0.0
        want = float(g[f"{name}_gnorm_net_{k}"])
        assert abs(gn - want) <= 1e-5 * max(want, 1e-12), k
        if f"{name}_g_net_{k}" in g.files:
            assert util.rel_err(p.grad.numpy(), g[f"{name}_g_net_{k}"]) < 1e-5, k
