"""The drop-in boundary against the reference tree itself: with this repo's
packages in front of /root/reference on sys.path, the import blocks of the reference's train_gui.py and render_mesh.py
must resolve every name they take from the packages this repo replaces, and ``install_into_reference()`` must yield
SUBCLASSES of the reference's own ControlNodeWarp / DeformNetwork that inherit everything outside the fast path.
Runs in a subprocess (tests/helpers/dropin_probe.py) because it stubs absent third-party GUI / mesh libraries.
Skipped where the reference tree is absent (the GPU box)."""
import json
import os
import subprocess
import sys

import pytest
import torch

import util

REF = "/root/reference"
PKG = os.path.join(util.ROOT, "dynamic-2dgs_b200")


@pytest.fixture(scope="module")
def probe():
    if not os.path.isdir(os.path.join(REF, "gaussian_renderer")):
        pytest.skip("reference tree not present")
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "helpers", "dropin_probe.py"), REF],
                       capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("PROBE_JSON ")]
    assert line, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(line[0][len("PROBE_JSON "):])


def test_reference_import_blocks_resolve(probe):
    tg, rm = probe["train_gui"], probe["render_mesh"]
    # train_gui.py:18 / :20, render_mesh.py:13,16,22
    for n in ("render", "network_gui", "render_flow"):
        assert tg["names"][n]["file"].startswith(PKG), (n, tg["names"][n])
    for n in ("Scene", "GaussianModel", "DeformModel"):
        assert tg["names"][n]["file"].startswith(REF), (n, tg["names"][n])
    assert rm["names"]["render"]["file"].startswith(PKG)
    assert rm["names"]["GaussianModel"]["module"] == "scene.gaussian_model"      # `from gaussian_renderer import GaussianModel`
    # whatever failed to import failed inside an absent third-party library, never inside a package this repo replaces
    ours = ("gaussian_renderer", "diff_surfel_rasterization", "diff_gaussian_rasterization", "simple_knn", "d2gs_b200",
            "scene.deform_model", "utils.time_utils")
    for stmt, err in {**tg["errors"], **rm["errors"]}.items():
        assert not any(o in stmt for o in ours), (stmt, err)
    assert "pytorch3d" not in probe["stubbed"] or probe["pytorch3d_shim"]


def test_bound_classes_subclass_the_reference(probe):
    b = probe["bound"]
    assert b["node_is_subclass"] and b["mlp_is_subclass"] and b["forward_is_fast"] and b["cal_nn_weight_is_fast"]
    assert all(b["inherited"].values()), b["inherited"]          # arap_loss, densify, as_gaussians, gs_* state ... are the reference's
    assert b["render_is_ours"] == "gaussian_renderer" and b["DeformModel_dict_is_patched"]
    assert b["network_class"][:2] == ["DeformNetwork", "_FusedNetworkMixin"]
    # calls outside the fast path (CPU tensors here) run the reference's code and keep its output contract
    assert b["cpu_forward_keys"] == ["d_color", "d_opacity", "d_rotation", "d_scaling", "d_xyz"]
    assert b["cpu_forward_shapes"] == [[200, 3], [200, 4], [200, 2]] and b["reg_loss_is_tensor"]
    # the stand-alone class's plain-torch fallbacks agree with the reference's code on the same state
    assert b["standalone_knn_matches_reference"]
    a_ref, a_alone = b["arap"]
    assert a_ref > 1e-6 and abs(a_ref - a_alone) <= 1e-4 * a_ref, b["arap"]


def test_standalone_class_keeps_the_trainer_contract():
    """Without the reference tree: state-dict keys incl. the gs_* entries of the warm-up node Gaussians round-trip,
    (M*T)-row time batches go through node_deform, the regularisers run, densify() is a no-op when disabled."""
    from d2gs_b200 import deform as dfm
    torch.manual_seed(1)
    cn = dfm.ControlNodeWarp(is_blender=True, node_num=40, K=4, hyper_dim=8, local_frame=True, with_arap_loss=True)
    sd = cn.state_dict()
    assert {"nodes", "_node_radius", "_node_weight", "inited", "network.linear.0.weight", "network.local_rotation.weight"} <= set(sd)
    sd["gs__xyz"] = torch.randn(40, 3); sd["gs__opacity"] = torch.randn(40, 1)
    sd["nodes"] = torch.randn(48, 11); sd["_node_radius"] = torch.randn(48); sd["_node_weight"] = torch.zeros(48, 1)   # densified checkpoint
    cn.load_state_dict(sd)
    assert cn.node_num == 48
    out = cn.state_dict()
    assert torch.equal(out["gs__xyz"], sd["gs__xyz"]) and torch.equal(out["gs__opacity"], sd["gs__opacity"])
    with torch.no_grad():
        cn.network.gaussian_warp.weight.mul_(3e3)
    t = torch.rand(48, 3, 1)
    v = cn.node_deform(t)
    assert v["d_xyz"].shape == (48, 3, 3) and v["local_rotation"].shape == (48, 3, 4)
    for fn in (cn.arap_loss, cn.elastic_loss, cn.acc_loss):
        l = fn(t=torch.tensor(0.5))
        assert l.dim() == 0 and torch.isfinite(l) and l.requires_grad
    assert cn.arap_loss(t=torch.tensor(0.5)) > 0
    assert cn.densify(max_grad=1e-3, optimizer=None, x=None, x_grad=None) is None
    assert dfm.landmark_interpolate([1e-4, 1e-4, 1e-5, 1e-5, 0], [0, 5000, 10000, 20000, 20001], 7500) == pytest.approx(10 ** -4.5)
    assert dfm.landmark_interpolate([0], [0], 5) == 0
    # softmax kernel / external node set: the plain-torch branch
    w, d, i = cn.cal_nn_weight(torch.randn(30, 3), K=5, nodes=torch.randn(20, 3), gs_kernel=False, temperature=0.1)
    assert w.shape == (30, 5) and torch.allclose(w.sum(-1), torch.ones(30)) and bool((d[:, 1:] >= d[:, :-1]).all())


def test_fused_path_is_not_taken_for_models_that_override_getters():
    """StandardGaussianModel-like objects (get_scaling overridden / `all_the_same`) must use the eager sequence."""
    import gaussian_renderer as gr
    from d2gs_b200 import model as mdl, synthetic as syn
    sc = syn.make_scene(16, 3)
    plain = mdl.SurfelModel(sc, device="cpu")
    assert gr._plain_getters(plain)

    class Iso(mdl.SurfelModel):
        @property
        def get_scaling(self):
            return torch.exp(self._scaling.mean(dim=1, keepdim=True).expand_as(self._scaling))
    assert not gr._plain_getters(Iso(sc, device="cpu"))
    flagged = mdl.SurfelModel(sc, device="cpu")
    flagged.all_the_same = False
    assert not gr._plain_getters(flagged)


def test_pytorch3d_stand_in_semantics():
    from d2gs_b200 import pytorch3d_shim as sh
    torch.manual_seed(0)
    a, b = torch.randn(1, 50, 3), torch.randn(1, 30, 3)
    b[0, 7] = b[0, 3]                                   # exact tie: lower index first
    r = sh.knn_points(a, b, K=4, return_nn=True)
    d = ((a[0][:, None] - b[0][None]) ** 2).sum(-1)
    want = torch.sort(d, dim=1, stable=True)
    assert torch.equal(r.idx[0], want.indices[:, :4]) and torch.allclose(r.dists[0], want.values[:, :4])
    assert torch.equal(r.knn[0], b[0][r.idx[0]])
    q = sh.ball_query(a, b, K=3, radius=1.0, return_nn=False)
    for i in range(50):
        hits = torch.nonzero(d[i] < 1.0).flatten()[:3]
        assert q.idx[0, i, :len(hits)].tolist() == hits.tolist() and (q.idx[0, i, len(hits):] == -1).all()


def test_permute_surfels_reorders_parameters_state_and_buffers():
    """layout.permute_surfels_: one permutation applied to every per-surfel tensor (parameters keep their identity, so the
    optimiser state stays attached), optimiser moments and plain buffers; tensors of other sizes are left alone."""
    import torch
    from d2gs_b200 import layout
    P = 37
    g = torch.Generator().manual_seed(0)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._xyz = torch.nn.Parameter(torch.randn(P, 3, generator=g))
            self._features_rest = torch.nn.Parameter(torch.randn(P, 15, 3, generator=g))
            self.other = torch.nn.Parameter(torch.randn(5, 3, generator=g))
            self.register_buffer("denom", torch.arange(P, dtype=torch.float32)[:, None])
            self.max_radii2D = torch.arange(P, dtype=torch.float32) * 2
    m = M()
    opt = torch.optim.Adam([{"params": [m._xyz], "lr": 1e-2}, {"params": [m._features_rest], "lr": 1e-2}, {"params": [m.other], "lr": 1e-2}])
    (m._xyz.sum() * 2 + (m._features_rest ** 2).sum() + m.other.sum()).backward()
    opt.step()
    before = {k: v.detach().clone() for k, v in dict(xyz=m._xyz, rest=m._features_rest, other=m.other, denom=m.denom, rad=m.max_radii2D,
                                                      m1=opt.state[m._xyz]["exp_avg"], v2=opt.state[m._features_rest]["exp_avg_sq"]).items()}
    ids = (id(m._xyz), id(m._features_rest))
    perm = torch.randperm(P, generator=g)
    n = layout.permute_surfels_(m, perm, optimizers=[opt])
    assert n == 4 + 4          # 2 parameters + buffer + attribute, 2 x (exp_avg, exp_avg_sq)
    assert (id(m._xyz), id(m._features_rest)) == ids and m._xyz.grad is None
    assert torch.equal(m._xyz, before["xyz"][perm]) and torch.equal(m._features_rest, before["rest"][perm])
    assert torch.equal(m.denom, before["denom"][perm]) and torch.equal(m.max_radii2D, before["rad"][perm])
    assert torch.equal(opt.state[m._xyz]["exp_avg"], before["m1"][perm]) and torch.equal(opt.state[m._features_rest]["exp_avg_sq"], before["v2"][perm])
    assert torch.equal(m.other, before["other"])
    with pytest.raises(ValueError):
        layout.permute_surfels_(m, torch.zeros(P, dtype=torch.long))


def test_deferred_count_check_filters_by_device_for_both_rasterizers():
    """raster.check_deferred_counts(device): the surfel rasterizer tracks counts under (device, P, W, H), the depth/alpha
    side rasterizer under ("gs3d", device, P, W, H) in the same table; a device filter must reach both, and "cuda"
    without an index means every device."""
    from d2gs_b200 import raster

    class FakeTrack:
        def __init__(self): self.polled = 0
        def poll(self, key, block=True): self.polled += 1
        def raise_if_overflowed(self): pass

    saved = dict(raster._TRACK)
    raster._TRACK.clear()
    try:
        keys = [(0, 10, 8, 8), (1, 10, 8, 8), ("gs3d", 0, 10, 8, 8), ("gs3d", 1, 10, 8, 8)]
        for k in keys:
            raster._TRACK[k] = FakeTrack()
        raster.check_deferred_counts(torch.device("cuda", 0))
        assert [raster._TRACK[k].polled for k in keys] == [1, 0, 1, 0]
        raster.check_deferred_counts("cuda:1")
        assert [raster._TRACK[k].polled for k in keys] == [1, 1, 1, 1]
        raster.check_deferred_counts("cuda")
        raster.check_deferred_counts(None)
        assert [raster._TRACK[k].polled for k in keys] == [3, 3, 3, 3]
    finally:
        raster._TRACK.clear()
        raster._TRACK.update(saved)
