"""GPU: the fused pieces of render() — raw-parameter rasterizer mode (activations + deltas in-kernel) and the fused
image-space epilogue — against the reference's eager op sequence (gaussian_renderer/__init__.py:83-99,172-207,
utils/point_utils.py:9-38) evaluated with torch on the same device."""
import math

import numpy as np
import pytest
import torch

import util
from oracle import reference_pipeline as rp

pytestmark = pytest.mark.gpu


def _scene(dev, cfg="T1", cam=5):
    from d2gs_b200 import model as mdl, synthetic as syn
    c = syn.CONFIGS[cfg]
    sc = syn.make_scene(c["P"], c["seed"], c["s_med"], n_nodes=0)
    cam = mdl.ViewCamera(syn.fibonacci_cameras(8, c["W"], c["H"])[cam], dev)
    return sc, cam, mdl


def test_raw_parameter_mode_equals_eager_glue(cuda_device):
    import diff_surfel_rasterization as ours
    from d2gs_b200 import raster
    dev = cuda_device
    sc, cam, mdl = _scene(dev)
    g = torch.Generator().manual_seed(0)
    P = sc.P
    d_xyz = (0.01 * torch.randn(P, 3, generator=g)).to(dev).requires_grad_(True)
    d_rot = (0.05 * torch.randn(P, 4, generator=g)).to(dev).requires_grad_(True)
    d_sc = (0.001 * torch.rand(P, 2, generator=g)).to(dev).requires_grad_(True)
    kw = dict(bg=(0.3, 0.1, 0.2), viewmatrix=cam.world_view_transform.cpu().numpy(), projmatrix=cam.full_proj_transform.cpu().numpy(),
              campos=cam.camera_center.cpu().numpy(), tanfovx=math.tan(cam.FoVx / 2), tanfovy=math.tan(cam.FoVy / 2),
              image_height=cam.image_height, image_width=cam.image_width, sh_degree=3)
    rs = util.settings_for(ours, kw, dev)
    gc, go = (torch.as_tensor(a, device=dev) for a in util.upstream_grads(cam.image_height, cam.image_width, seed=9))
    res = {}
    for mode in ("eager", "raw"):
        pc = mdl.SurfelModel(sc, dev)
        for t in (d_xyz, d_rot, d_sc):
            t.grad = None
        m2d = torch.zeros_like(pc._xyz, requires_grad=True)
        if mode == "eager":
            color, radii, allmap = raster.rasterize_surfels(pc.get_xyz + d_xyz, m2d, pc._features_dc, None, pc.get_opacity,
                                                            pc.get_scaling + d_sc, pc.get_rotation_bias(d_rot), None, rs,
                                                            sh_rest=pc._features_rest)
        else:
            color, radii, allmap = raster.rasterize_surfels_raw(pc._xyz, d_xyz, pc._scaling, d_sc, pc._rotation, d_rot,
                                                                pc._opacity, m2d, pc._features_dc, pc._features_rest, None, rs)
        ((color * gc).sum() + (allmap * go).sum()).backward()
        torch.cuda.synchronize()
        res[mode] = dict(color=color.detach(), radii=radii, allmap=allmap.detach(),
                         grads={n: p.grad.clone() for n, p in pc.named_parameters() if p.grad is not None},
                         d=(d_xyz.grad.clone(), d_rot.grad.clone(), d_sc.grad.clone()), m2d=m2d.grad.clone())
    e, r = res["eager"], res["raw"]
    # exp / sigmoid / F.normalize are restated op for op: identical tile decisions, images equal to fp32 round-off
    assert torch.equal(e["radii"], r["radii"])
    assert util.rel_err(r["color"].cpu().numpy(), e["color"].cpu().numpy()) < 1e-6
    assert util.rel_err(r["allmap"][:6].cpu().numpy(), e["allmap"][:6].cpu().numpy()) < 1e-6
    assert util.rel_err(r["allmap"][6].cpu().numpy(), e["allmap"][6].cpu().numpy()) < 1e-3   # distortion: variance-like, ill-conditioned
    assert set(e["grads"]) == set(r["grads"])
    for n in e["grads"]:
        assert util.rel_err(r["grads"][n].cpu().numpy(), e["grads"][n].cpu().numpy()) < 2e-5, n
    for a, b, n in zip(r["d"], e["d"], ("d_xyz", "d_rot", "d_sc")):
        assert util.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, n
    assert util.rel_err(r["m2d"].cpu().numpy(), e["m2d"].cpu().numpy()) < 2e-5
    # deltas given as python 0.0 (train_gui.py:265,480) -> same as zero tensors
    pc = mdl.SurfelModel(sc, dev)
    c0, _, a0 = raster.rasterize_surfels_raw(pc._xyz, None, pc._scaling, None, pc._rotation, None, pc._opacity,
                                             torch.zeros_like(pc._xyz), pc._features_dc, pc._features_rest, None, rs)
    z = torch.zeros
    c1, _, a1 = raster.rasterize_surfels_raw(pc._xyz, z(P, 3, device=dev), pc._scaling, z(P, 2, device=dev), pc._rotation,
                                             z(P, 4, device=dev), pc._opacity, torch.zeros_like(pc._xyz), pc._features_dc,
                                             pc._features_rest, None, rs)
    assert torch.equal(c0, c1) and torch.equal(a0, a1)


def test_fused_epilogue_equals_eager(cuda_device):
    from d2gs_b200 import epilogue
    dev = cuda_device
    _, cam, _ = _scene(dev)
    H, W = cam.image_height, cam.image_width
    g = torch.Generator().manual_seed(3)
    allmap = torch.rand((8, H, W), generator=g)
    allmap[5] = 2.0 + allmap[5] * 3 + 0.3 * torch.sin(torch.arange(W)[None] / 7.0) * torch.cos(torch.arange(H)[:, None] / 5.0)
    allmap[5, 10:14, 20:24] = float("nan"); allmap[5, 30, 40] = float("inf"); allmap[5, :4, :] = 0; allmap[1, :4, :] = 0
    allmap = allmap.to(dev)
    ups = [torch.randn(s, generator=g).to(dev) for s in ((1, H, W), (3, H, W), (1, H, W), (1, H, W), (3, H, W), (3, H, W))]

    def eager(A):
        alpha = A[1:2]
        rn = (A[2:5].permute(1, 2, 0) @ (cam.world_view_transform[:3, :3].T)).permute(2, 0, 1)
        med = torch.nan_to_num(A[5:6], 0, 0)
        exp = torch.nan_to_num(A[0:1] / alpha, 0, 0)
        dist = A[6:7]
        sd = exp * 0 + 1 * med
        sn, sp = rp.depth_to_normal(cam, sd)
        sn = sn.permute(2, 0, 1) * alpha.detach()
        return alpha, rn, dist, sd, sn, sp.permute(2, 0, 1)

    A1 = allmap.clone().requires_grad_(True)
    o1 = eager(A1)
    sum((o * u).sum() for o, u in zip(o1, ups)).backward()
    A2 = allmap.clone().requires_grad_(True)
    o2 = epilogue.render_epilogue(A2, cam)
    sum((o * u).sum() for o, u in zip(o2, ups)).backward()
    torch.cuda.synchronize()
    names = ("alpha", "rend_normal", "rend_dist", "depth", "surf_normal", "surf_point")
    for n, a, b in zip(names, o2, o1):
        assert a.shape == b.shape, n
        a_, b_ = a.detach().cpu().numpy(), b.detach().cpu().numpy()
        bad = ~np.isclose(a_, b_, rtol=2e-4, atol=2e-5)
        # normalising a cross product of nearly parallel finite differences is ill-conditioned: allow isolated pixels
        assert bad.mean() < 1e-4 and np.abs(a_ - b_).max() < 1e-3, (n, bad.sum(), np.abs(a_ - b_).max())
    ga, gb = A2.grad.cpu().numpy(), A1.grad.cpu().numpy()
    # eager yields 0/0 = NaN on planes 0/1 where alpha == 0 (the unused expected-depth branch); those pixels have no
    # contributors, so the rasterizer backward never reads them.  Compare where eager is finite; ours must be finite everywhere.
    assert np.isfinite(ga).all() and not ga[0].any()
    for pl in range(1, 8):
        sel = np.isfinite(gb[pl])
        assert sel.mean() > 0.9
        assert util.rel_err(ga[pl][sel], gb[pl][sel]) < 2e-4, pl
    nanpix = ~np.isfinite(gb[1])
    assert np.allclose(ga[1][nanpix], ups[0].cpu().numpy()[0][nanpix])     # there our alpha gradient is simply g_alpha
    # outputs that are not used downstream give None grads: must be accepted
    A3 = allmap.clone().requires_grad_(True)
    o3 = epilogue.render_epilogue(A3, cam)
    (o3[1] * ups[1]).sum().backward()
    assert torch.isfinite(A3.grad).all() and not A3.grad[5].any()


def test_step_replays_as_cuda_graph(cuda_device):
    """DeformModel.step + render + fused loss + backward recorded ONCE as a CUDA graph through the public API (the
    rasterizer switches to its deferred-count mode: nothing synchronises), then replayed for a DIFFERENT view and time
    written into the static camera tensors: images and gradients must equal the eager result for that view."""
    from d2gs_b200 import deform as dfm, model as mdl, raster, synthetic as syn
    from d2gs_b200.loss import surfel_loss
    from gaussian_renderer import render
    dev = cuda_device
    c = syn.CONFIGS["T1"]
    sc = syn.make_scene(c["P"], c["seed"], c["s_med"], n_nodes=c["n_nodes"], hyper_dim=8)
    cams = syn.fibonacci_cameras(8, c["W"], c["H"])
    pc, pipe = mdl.SurfelModel(sc, dev), mdl.PipelineParams()
    torch.manual_seed(0)
    dm = dfm.DeformModel(deform_type="node", is_blender=True, K=c["K"], hyper_dim=8, node_num=c["n_nodes"], local_frame=True)
    with torch.no_grad():
        dm.deform.nodes.copy_(torch.as_tensor(sc.nodes, device=dev))
        dm.deform._node_radius.copy_(torch.as_tensor(sc.node_radius, device=dev))
        dm.deform.network.gaussian_warp.weight.mul_(1e3)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    gt = torch.rand((3, c["H"], c["W"]), generator=torch.Generator().manual_seed(1)).to(dev)
    params = [pc._xyz, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling, pc._rotation, dm.deform.network.gaussian_warp.weight]
    cam = mdl.ViewCamera(cams[1], dev)          # the static camera the graph reads

    def set_view(i):
        v = mdl.ViewCamera(cams[i], dev)
        for n in ("world_view_transform", "full_proj_transform", "camera_center", "fid"):
            getattr(cam, n).copy_(getattr(v, n))

    def step():
        for p in params:
            p.grad = None
        d = dm.step(pc.get_xyz.detach(), dm.deform.expand_time(cam.fid), feature=pc.feature, motion_mask=pc.motion_mask)
        out = render(cam, pc, pipe, bg, d["d_xyz"], d["d_rotation"], d["d_scaling"])
        loss = surfel_loss(out["render"], gt, out["rend_normal"], out["surf_normal"], out["rend_dist"], 0.2, 0.02, 100.0)
        loss.backward()
        return out["render"], loss, [p.grad for p in params]

    for i in (1, 2, 3):       # eager frames: they establish the instance-count history the capture sizes its binning from
        set_view(i)
        step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        image, loss, grads = step()
    set_view(6)
    g.replay()
    torch.cuda.synchronize()
    got = [image.clone(), loss.clone()] + [x.clone() for x in grads]
    raster.check_deferred_counts(dev)
    del g
    raster.set_deferred_count(False)
    image_e, loss_e, grads_e = step()
    torch.cuda.synchronize()
    assert torch.equal(got[0], image_e)
    assert abs(float(got[1]) - float(loss_e)) <= 1e-6 * abs(float(loss_e))
    for a, b, p in zip(got[2:], grads_e, params):
        assert util.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, tuple(p.shape)
    assert float(got[0].std()) > 0.01


def test_storage_order_changes_nothing(cuda_device):
    """layout.permute_surfels_ (Morton-sorted storage): deform + render + backward + one FusedAdam step of a
    permuted model equal those of the original model up to the permutation — images to fp32 round-off (the blend order inside a
    tile is (depth, surfel id): exact depth ties are the only thing the ids can reorder), gradients and updated parameters
    row for row."""
    from d2gs_b200 import deform as dfm, layout, model as mdl, synthetic as syn
    from d2gs_b200.optim import FusedAdam
    from gaussian_renderer import render
    dev = cuda_device
    cfg = syn.CONFIGS["T1"]
    sc = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"], n_nodes=cfg["n_nodes"], hyper_dim=8)
    cam = mdl.ViewCamera(syn.fibonacci_cameras(8, cfg["W"], cfg["H"])[3], dev)
    g = torch.Generator().manual_seed(4)
    w = (torch.randn((3, cfg["H"], cfg["W"]), generator=g) / (cfg["H"] * cfg["W"])).to(dev)
    res = {}
    perm = None
    for mode in ("as_generated", "morton"):
        torch.manual_seed(0)
        pc = mdl.SurfelModel(sc, dev)
        dm = dfm.DeformModel(deform_type="node", is_blender=True, K=4, hyper_dim=8, node_num=cfg["n_nodes"], local_frame=True)
        with torch.no_grad():
            dm.deform.nodes.copy_(torch.as_tensor(sc.nodes, device=dev))
            dm.deform._node_radius.copy_(torch.as_tensor(sc.node_radius, device=dev))
            dm.deform.network.gaussian_warp.weight.mul_(1e3)
        opt = FusedAdam([{"params": [pc._xyz], "lr": 1e-3, "name": "xyz"}, {"params": [pc._features_rest], "lr": 1e-3, "name": "f_rest"}],
                        lr=0.0, eps=1e-15)
        if mode == "morton":
            perm = layout.morton_permutation(pc.get_xyz)
            n = layout.permute_surfels_(pc, perm, optimizers=[opt])
            assert n >= 8 and not torch.equal(perm, torch.arange(perm.numel(), device=perm.device))
        for step in range(2):            # second step: the optimiser state (first step's moments) is in play
            d = dm.step(pc.get_xyz.detach(), dm.deform.expand_time(cam.fid), feature=pc.feature, motion_mask=pc.motion_mask)
            out = render(cam, pc, mdl.PipelineParams(), torch.zeros(3, device=dev), d["d_xyz"], d["d_rotation"], d["d_scaling"])
            loss = (out["render"] * w).sum() + 1e-3 * out["rend_dist"].sum()
            for p_ in pc.parameters():
                p_.grad = None
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        res[mode] = dict(img=out["render"].detach().cpu().numpy(), radii=out["radii"].cpu().numpy(), xyz=pc._xyz.detach().cpu().numpy(),
                         g_rest=pc._features_rest.grad.cpu().numpy(), g_op=pc._opacity.grad.cpu().numpy())
    a, b, p = res["as_generated"], res["morton"], perm.cpu().numpy()
    assert np.array_equal(b["radii"], a["radii"][p])
    assert util.rel_err(b["img"], a["img"]) < 1e-5
    assert util.rel_err(b["xyz"], a["xyz"][p]) < 1e-6
    assert util.rel_err(b["g_rest"], a["g_rest"][p]) < 1e-4 and util.rel_err(b["g_op"], a["g_op"][p]) < 1e-4
