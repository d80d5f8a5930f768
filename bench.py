#!/usr/bin/env python
"""bench.py — fwd+bwd frames/s of the Dynamic-2DGS per-frame hot path (deform -> render -> loss -> backward).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference] [--config C3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric: "fwd+bwd frames/sec @300k surfels 800x800"): config C3 = 300k surfels + 512 control
nodes (K=4, hyper_dim 8, local_frame), SH degree 3, 800x800, 100 seeded views, synthetic data, random-init weights.
One "step" = DeformModel.step + render() + fixed synthetic loss + backward for ONE view per GPU; with N>1 views are
sharded over ranks (weak scaling) and the flat gradient bucket is all-reduced (NCCL) inside the timed region.

The single JSON line printed by rank 0 follows the driver contract (metric/value/unit/n_gpus/steps/warmup/
ms_per_step/higher_is_better/scaling/vs_baseline/dtype/data/config/clocks/e2e/gpu_launches) plus `roofline`
(dominant kernel, CUDA-event timed inside the timed region) and `cpu_baseline` (CPU oracle port on the host cores).
`--impl reference` times the reference pipeline: the UNMODIFIED reference CUDA rasterizer (oracle/_ref) inside the
reference's eager-torch op sequence (oracle/reference_pipeline.py); if the extension cannot be loaded it times the
CPU oracle port instead and says so.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dynamic-2dgs_b200"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from d2gs_b200 import synthetic as syn

METRIC = "fwd+bwd frames/sec @300k surfels 800x800"
UNIT = "frames/s"
N_VIEWS = 100
TRAIN_LAMBDAS = (0.2, 0.02, 1000.0)   # lambda_dssim (arguments/__init__.py), lambda_normal / lambda_dist after iteration 8000 (train_gui.py:292-293)
HEAD_SCALE = 1e3      # default head init is ~1e-5, scaled so the deformation is non-trivial


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, read through NVML in a background thread.

    (Spawning `nvidia-smi -lms` next to the benchmark was measured to slow every CUDA launch of this process by ~3x
    on these hosts — the tool re-enumerates the devices under a driver lock on each sample — so the same counters are
    read in-process; `nvidia-smi` is only the fallback when pynvml is missing.)"""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.gpu, self.period, self.samples, self.stop_flag, self.thread = gpu_index, period_s, [], False, None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(sm), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.h is None:
            return self._smi_once()
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            self._loop_once()
        sm = [s for s, _ in self.samples]
        bits = 0
        for _, r in self.samples:
            bits |= r
        reasons = sorted(n for b, n in self.REASONS.items() if bits & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(sm), "source": "nvml"}

    def _loop_once(self):
        self.stop_flag = True
        try:
            nv = self.nv
            self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), 0))
        except Exception:
            pass

    def _smi_once(self) -> dict:
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                  str(self.gpu)], capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi (after the run)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}


# ----------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------
def loss_weights(H, W, device, seed=99):
    g = torch.Generator().manual_seed(seed)
    mk = lambda c: (torch.randn((c, H, W), generator=g) / (H * W)).to(device)
    return {"render": mk(3), "alpha": mk(1), "rend_normal": mk(3), "rend_dist": mk(1), "depth": mk(1)}


class _SyntheticLoss(torch.autograd.Function):
    """loss = mean|render - gt| + sum_k <out_k, w_k> with as few eager launches as torch allows (both arms call it): one
    `dot` per weighted map forward, one scaled copy of the constant weight per map backward (round 1 spent 30 launches /
    ~0.15 ms of every step here — 11 % of our kernel time — in mul / sum / add / expand kernels)."""

    @staticmethod
    def forward(ctx, gt, w_render, w_alpha, w_normal, w_dist, w_depth, render, alpha, normal, dist, depth):
        diff = render - gt
        l1 = diff.abs().mean()
        dots = torch.stack([torch.dot(render.reshape(-1), w_render.reshape(-1)), torch.dot(alpha.reshape(-1), w_alpha.reshape(-1)),
                            torch.dot(normal.reshape(-1), w_normal.reshape(-1)), torch.dot(dist.reshape(-1), w_dist.reshape(-1)),
                            torch.dot(depth.reshape(-1), w_depth.reshape(-1))])
        # d loss / d render = w_render + sign(render - gt) / N: formed here, scaled by the upstream gradient in backward
        ctx.save_for_backward(torch.add(w_render, torch.sign(diff), alpha=1.0 / diff.numel()), w_alpha, w_normal, w_dist, w_depth)
        return l1 + dots.sum()

    @staticmethod
    def backward(ctx, g):
        c_render, w_alpha, w_normal, w_dist, w_depth = ctx.saved_tensors
        return (None,) * 6 + (c_render * g, w_alpha * g, w_normal * g, w_dist * g, w_depth * g)


def synthetic_loss(out, wts, gt):
    """Fixed seeded random-weighted sum over the render() outputs + an L1 term to the target image."""
    return _SyntheticLoss.apply(gt, wts["render"], wts["alpha"], wts["rend_normal"], wts["rend_dist"], wts["depth"],
                                out["render"], out["alpha"], out["rend_normal"], out["rend_dist"], out["depth"])


class Workload:
    """Device-resident state of one rank: surfel model, 100 cameras, loss weights, target image."""

    def __init__(self, cfg_name: str, device: torch.device, impl: str):
        from d2gs_b200 import model as mdl
        cfg = dict(syn.CONFIGS[cfg_name])
        self.cfg_name, self.cfg, self.device, self.impl = cfg_name, cfg, device, impl
        self.scene = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"], n_nodes=cfg["n_nodes"], hyper_dim=8)
        self.cams_np = syn.fibonacci_cameras(N_VIEWS, cfg["W"], cfg["H"])
        self.pc = mdl.SurfelModel(self.scene, device)
        self.cams = [mdl.ViewCamera(c, device, uid=i) for i, c in enumerate(self.cams_np)]
        self.pipe = mdl.PipelineParams()
        self.bg = torch.zeros(3, device=device)
        self.W, self.H, self.K = cfg["W"], cfg["H"], cfg["K"]
        self.wts = loss_weights(self.H, self.W, device)
        self.gt_host = torch.rand((3, self.H, self.W), generator=torch.Generator().manual_seed(5)).pin_memory()
        self.gt_dev = self.gt_host.to(device)
        self.use_deform = cfg["n_nodes"] > 0
        self.loss_kind = "synthetic"
        self.keep_stats = False
        self.deform_parameters = lambda: []


def build_deform_ours(wl: Workload):
    from d2gs_b200 import deform as dfm
    torch.manual_seed(1234)
    dm = dfm.DeformModel(deform_type="node", is_blender=True, K=wl.K, hyper_dim=8, node_num=wl.cfg["n_nodes"], local_frame=True,
                         with_arap_loss=False)
    cn = dm.deform
    with torch.no_grad():
        cn.nodes.copy_(torch.as_tensor(wl.scene.nodes, device=wl.device))
        cn._node_radius.copy_(torch.as_tensor(wl.scene.node_radius, device=wl.device))
        cn._node_weight.copy_(torch.as_tensor(wl.scene.node_weight, device=wl.device))
        for head in (cn.network.gaussian_warp, cn.network.gaussian_scaling, cn.network.gaussian_rotation, cn.network.local_rotation):
            head.weight.mul_(HEAD_SCALE)
    cn.train()
    wl.deform = dm
    wl.deform_parameters = lambda: [p for p in cn.parameters() if p.requires_grad]


def build_deform_reference(wl: Workload):
    from oracle import deform_oracle as do
    p = do.init_network_params(seed=1234, local_frame=True, head_scale=HEAD_SCALE)
    wl.net = {k: v.to(wl.device).requires_grad_(True) for k, v in p.items()}
    t = lambda a: torch.as_tensor(a, device=wl.device).clone().requires_grad_(True)
    wl.nodes, wl.node_radius, wl.node_weight = t(wl.scene.nodes), t(wl.scene.node_radius), t(wl.scene.node_weight)
    wl.deform_parameters = lambda: list(wl.net.values()) + [wl.nodes, wl.node_radius, wl.node_weight]


def step_ours(wl: Workload, cam, gt, gt_ready=None):
    from gaussian_renderer import render
    pc = wl.pc
    if wl.use_deform:
        t_in = wl.deform.deform.expand_time(cam.fid)
        d = wl.deform.step(pc.get_xyz.detach(), t_in, feature=pc.feature, motion_mask=pc.motion_mask)
        d_xyz, d_rot, d_scale = d["d_xyz"], d["d_rotation"], d["d_scaling"]
    else:
        d_xyz, d_rot, d_scale = 0.0, 0.0, 0.0
    out = render(cam, pc, wl.pipe, wl.bg, d_xyz, d_rot, d_scale)
    if wl.keep_stats:      # only what the training tail reads; holding `out` would keep the step's autograd graph alive
        wl.last_out = {"viewspace_points": out["viewspace_points"], "visibility_filter": out["visibility_filter"], "radii": out["radii"]}
    if gt_ready is not None:
        torch.cuda.current_stream().wait_event(gt_ready)     # the target image arrives on the copy stream
    if wl.loss_kind == "train":
        # the training loss of train_gui.py:292-313 (L1 + D-SSIM + normal consistency + distortion), fused
        from d2gs_b200 import loss as fl
        loss = fl.surfel_loss(out["render"], gt, out["rend_normal"], out["surf_normal"], out["rend_dist"], *TRAIN_LAMBDAS)
    else:
        loss = synthetic_loss(out, wl.wts, gt)
    loss.backward()
    return loss


class RefStageTimer:
    """CUDA-event stage timers of the REFERENCE arm: events around the eager deformation, around
    the two native calls of the reference extension (`_C.rasterize_gaussians{,_backward}`, patched on its module object)
    and around the whole step; "epilogue+loss fwd" and "autograd bwd outside the rasterizer" are the remainders."""

    def __init__(self, ref_mod):
        self.mod, self.on, self.ev, self.orig = ref_mod, False, [], {}

    def _wrap(self, name, stage):
        fn = getattr(self.mod._C, name)
        self.orig[name] = fn

        def timed(*a, **k):
            if not self.on:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.ev.append((stage, e0, e1))
            return r
        setattr(self.mod._C, name, timed)

    def install(self):
        self._wrap("rasterize_gaussians", "raster_fwd")
        self._wrap("rasterize_gaussians_backward", "raster_bwd")

    def mark(self, stage):
        """Context manager: events around a python-level section."""
        timer = self

        class _M:
            def __enter__(self_):
                if timer.on:
                    self_.e0 = torch.cuda.Event(enable_timing=True); self_.e0.record()

            def __exit__(self_, *exc):
                if timer.on:
                    e1 = torch.cuda.Event(enable_timing=True); e1.record()
                    timer.ev.append((stage, self_.e0, e1))
        return _M()

    def collect(self, steps):
        torch.cuda.synchronize()
        acc = {}
        for stage, e0, e1 in self.ev:
            acc[stage] = acc.get(stage, 0.0) + e0.elapsed_time(e1)
        self.ev = []
        out = {k: v / steps for k, v in acc.items()}
        if "step" in out:
            fwd_rest = out.get("forward", 0.0) - out.get("deform_fwd", 0.0) - out.get("raster_fwd", 0.0)
            out["epilogue_loss_fwd"] = fwd_rest
            out["autograd_bwd_outside_rasterizer"] = out.get("backward", 0.0) - out.get("raster_bwd", 0.0)
        return out


class _Null:
    def __enter__(self): return self
    def __exit__(self, *a): return False


def step_reference(wl: Workload, cam, gt, gt_ready=None):
    from oracle import reference_pipeline as rp
    pc = wl.pc
    tm = getattr(wl, "ref_timer", None)
    mark = tm.mark if tm is not None else (lambda s_: _Null())
    with mark("forward"):
        with mark("deform_fwd"):
            if wl.use_deform:
                d = rp.deform_reference(wl.net, wl.nodes, wl.node_radius, wl.node_weight, pc.get_xyz.detach(), cam.fid, pc.feature,
                                        pc.motion_mask, wl.K, 8, local_frame=True)
                d_xyz, d_rot, d_scale = d["d_xyz"], d["d_rotation"], d["d_scaling"]
            else:
                d_xyz, d_rot, d_scale = 0.0, 0.0, 0.0
        out = rp.render_reference(wl.ref_mod, cam, pc, wl.bg, d_xyz, d_rot, d_scale)
        if wl.keep_stats:
            wl.last_out = {"viewspace_points": out["viewspace_points"], "visibility_filter": out["visibility_filter"], "radii": out["radii"]}
        if gt_ready is not None:
            torch.cuda.current_stream().wait_event(gt_ready)
        if wl.loss_kind == "train":
            from oracle import loss_oracle as lo      # eager restatement of utils/loss_utils.py + train_gui.py:292-313
            loss = lo.surfel_loss(out["render"], gt, out["rend_normal"], out["surf_normal"], out["rend_dist"], *TRAIN_LAMBDAS)[0]
        else:
            loss = synthetic_loss(out, wl.wts, gt)
    with mark("backward"):
        loss.backward()
    return loss


# ----------------------------------------------------------------------------------------------------------------
# rasterizer-only comparison (north_star: ">= 3x the reference diff-surfel-rasterization fwd+bwd at 300 k / 800x800")
# ----------------------------------------------------------------------------------------------------------------
def raster_only(cfg_name: str, device, impl: str, ref_mod=None, n_views: int = 10, reps: int = 3) -> dict:
    """fwd+bwd of the rasterizer OP ALONE through the reference's own module interface
    (`GaussianRasterizer(settings)(means3D, means2D, shs, opacities, scales, rotations)` -> colour + 8 planes, loss = seeded
    random-weighted sum over all 11 planes, `.backward()`): the scene of `cfg_name` with the deformation off (activated
    parameters as render() would pass them), `n_views` of the bench's cameras, `reps` passes, eager launches with the
    op's default settings in both arms (ours: per-tile binning, synchronous instance count — exactly what a trainer that
    only swaps the package gets).  CUDA events over the whole loop, one synchronisation at the end."""
    if impl == "ours":
        import diff_surfel_rasterization as mod
    else:
        mod = ref_mod
    cfg = syn.CONFIGS[cfg_name]
    sc = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"])
    act = syn.activated(sc)
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    ins = {k: T(v).requires_grad_(True) for k, v in act.items()}
    cams = syn.fibonacci_cameras(N_VIEWS, cfg["W"], cfg["H"])
    views = [(i * 7 + 3) % N_VIEWS for i in range(n_views)]
    g = torch.Generator().manual_seed(17)
    gc = (torch.randn((3, cfg["H"], cfg["W"]), generator=g) / (cfg["H"] * cfg["W"])).to(device)
    go = (torch.randn((8, cfg["H"], cfg["W"]), generator=g) / (cfg["H"] * cfg["W"])).to(device)
    bg = torch.zeros(3, device=device)
    settings = []
    for v in views:
        c = cams[v]
        settings.append(mod.GaussianRasterizationSettings(
            image_height=c.image_height, image_width=c.image_width, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
            viewmatrix=T(c.world_view_transform), projmatrix=T(c.full_proj_transform), sh_degree=3, campos=T(c.camera_center),
            prefiltered=False, debug=False))

    def one(rs):
        for t in ins.values():
            t.grad = None
        m2d = torch.zeros_like(ins["means3D"], requires_grad=True)
        color, radii, allmap = mod.GaussianRasterizer(rs)(means3D=ins["means3D"], means2D=m2d, shs=ins["shs"], opacities=ins["opacities"],
                                                          scales=ins["scales"], rotations=ins["rotations"])
        ((color * gc).sum() + (allmap * go).sum()).backward()

    for rs in settings[:3]:
        one(rs)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for rs in settings:
            one(rs)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / (reps * len(settings))
    return {"ms": ms, "frames_per_s": 1e3 / ms, "workload": f"{cfg_name} scene, deformation off: {cfg['P']} surfels, SH3, {cfg['W']}x{cfg['H']}, "
            f"{len(settings)} views x {reps}, rasterizer op fwd+bwd through GaussianRasterizer (eager, default settings)"}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline (oracle port on host cores) — a reported baseline, bounded sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg_name: str, frames: int = 1) -> dict:
    from oracle import surfel_oracle as so
    from oracle import deform_oracle as do
    cfg = syn.CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    so.set_num_threads(cores)
    torch.set_num_threads(cores)
    sc = syn.make_scene(cfg["P"], cfg["seed"], cfg["s_med"], n_nodes=cfg["n_nodes"], hyper_dim=8)
    cams = syn.fibonacci_cameras(N_VIEWS, cfg["W"], cfg["H"])
    rng = np.random.default_rng(0)
    gc = (rng.normal(size=(3, cfg["H"], cfg["W"])) / (cfg["H"] * cfg["W"])).astype(np.float32)
    go = (rng.normal(size=(8, cfg["H"], cfg["W"])) / (cfg["H"] * cfg["W"])).astype(np.float32)
    tt = lambda a: torch.as_tensor(a)
    xyz, scaling, rotation, opacity = tt(sc.xyz), tt(sc.scaling).requires_grad_(True), tt(sc.rotation).requires_grad_(True), tt(sc.opacity).requires_grad_(True)
    feature = tt(sc.feature).requires_grad_(True)
    xyz_p = xyz.clone().requires_grad_(True)
    if cfg["n_nodes"]:
        net = {k: v.requires_grad_(True) for k, v in do.init_network_params(seed=1234, local_frame=True, head_scale=HEAD_SCALE).items()}
        nodes, nrad, nw = tt(sc.nodes).requires_grad_(True), tt(sc.node_radius).requires_grad_(True), tt(sc.node_weight).requires_grad_(True)
    shs = np.concatenate([sc.features_dc, sc.features_rest], 1)
    t0 = time.perf_counter()
    R = 0
    for f in range(frames):
        cam = cams[(7 * f) % N_VIEWS]
        if cfg["n_nodes"]:
            t = torch.full((nodes.shape[0], 1), cam.fid)
            d = do.control_node_warp_forward(net, nodes, nrad, nw, xyz, t, feature, torch.ones(xyz.shape[0], 1), cfg["K"], 8,
                                             local_frame=True, knn_mode="mm")
            means3D, opac, scales, rot = do.render_glue_pre(xyz_p, scaling, rotation, opacity, d["d_xyz"], d["d_rotation"], d["d_scaling"])
        else:
            z = torch.zeros(())
            means3D, opac, scales, rot = do.render_glue_pre(xyz_p, scaling, rotation, opacity, z, z, z)
        st = so.forward(bg=np.zeros(3, np.float32), means3D=means3D.detach().numpy(), opacities=opac.detach().numpy(),
                        scales=scales.detach().numpy(), rotations=rot.detach().numpy(), shs=shs, sh_degree=3,
                        viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform, campos=cam.camera_center,
                        tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, image_height=cam.image_height, image_width=cam.image_width)
        g = so.backward(st, gc, go)
        R = st.num_rendered
        torch.autograd.backward([means3D, opac, scales, rot],
                                [tt(g["dL_dmeans3D"]), tt(g["dL_dopacity"]), tt(g["dL_dscales"]), tt(g["dL_drotations"])])
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{frames} frame(s) of {cfg_name} (deform in torch-CPU + C++/OpenMP oracle rasterizer fwd+bwd, fp32, R={R}), {dt:.2f} s"}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic bytes
# ----------------------------------------------------------------------------------------------------------------
def stage_bytes(P, R, HW, n_pass):
    return {
        "preprocess_fwd": 232 * P + 80 * P,
        "duplicate": 12 * R + 16 * P,
        "sort": 24 * R * n_pass,
        "blend_fwd": 96 * R + 64 * HW,            # 96-B projected record (incl. cull box) per instance + 16 planes
        "blend_bwd": 96 * R + 64 * HW + 80 * P,   # + the 80-B per-surfel gradient record
        "preprocess_bwd": (232 + 76 + 36 + 3) * P + 244 * P,
        "deform_fwd": (12 + 32) * P + 36 * P,
        "deform_bwd": (12 + 32) * P + 36 * P + 32 * P,
        "epilogue_fwd": (32 + 56) * HW,
        "epilogue_bwd": (32 + 56) * HW,
    }


def frame_bytes(P, R, HW, deform):
    return 979 * P + 316 * R + 128 * HW + (156 * P if deform else 0)


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--loss", choices=["synthetic", "train"], default="synthetic",
                    help="synthetic: seeded random-weighted sum + L1 (SURVEY 8(d), the headline); train: the reference's training loss "
                         "(L1 + D-SSIM + normal + distortion), fused kernel in our arm, eager torch in the reference arm")
    ap.add_argument("--train", action="store_true",
                    help="C4-style training step (train_gui.py:292-432): --loss train + densification statistics (all-reduced across "
                         "ranks) + max-radii bookkeeping + the optimiser steps of both parameter sets, all inside the timed step; "
                         "ours: FusedAdam (1 launch per optimiser), reference arm: torch.optim.Adam as the reference runs it")
    ap.add_argument("--no-raster-only", action="store_true", help="skip the rasterizer-only comparison block")
    ap.add_argument("--early-allreduce", type=int, default=1, help="1 (default): start the all-reduce of the surfel-table gradients right after the rasterizer backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-count", action="store_true",
                    help="read num_rendered back every forward like the reference (one stream synchronisation per frame) instead of "
                         "the deferred-count binning mode")
    ap.add_argument("--graph", choices=["auto", "on", "off"], default="auto",
                    help="replay the step (deform + render + loss + backward + all-reduce) as ONE CUDA graph captured through the public API; "
                         "auto = fall back to eager launches if capture fails.  The reference arm is always eager (it synchronises inside).")
    ap.add_argument("--tile-sort", type=int, default=None, choices=[0, 1],
                    help="binning: 1 = per-tile buckets + segmented sort, 0 = global radix sort (same lists); default: the library's")
    ap.add_argument("--no-clocks", action="store_true", help="diagnosis only: do not sample clocks during the timed region")
    ap.add_argument("--view-schedule", choices=["balanced", "roundrobin"], default="roundrobin",
                    help="N > 1 only: roundrobin (default) = view (step * N + rank) mod 100, the same order as at N = 1; balanced = the views "
                         "of one step (one per rank) are chosen to cost the same, so no rank waits for a slower one (every view still once "
                         "per epoch; measured: 1.737 vs 1.749 ms at N = 8 — the views of this scene cost alike)")
    ap.add_argument("--storage", choices=["morton", "random"], default="morton",
                    help="storage order of the surfel tables (both arms): morton = sorted along a Morton curve once at set-up with "
                         "d2gs_b200.layout.permute_surfels_, as a trainer does after densification; random = as generated")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.train:
        args.loss = "train"

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1

    if args.impl == "reference" and rank != 0:
        return 0   # the reference is single-GPU; rank 0 alone measures it

    if not torch.cuda.is_available():
        if args.impl == "reference":
            cb = cpu_baseline(args.config)
            print(json.dumps({"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0, "steps": 1, "warmup": 0,
                              "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "impl": "reference",
                              "cpu_baseline": dict(cb), "config": {"workload": args.config},
                              "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU oracle)")

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if distributed and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    else:
        dist = None

    from d2gs_b200 import model as mdl
    wl = Workload(args.config, device, args.impl)
    wl.loss_kind = args.loss
    wl.keep_stats = bool(args.train)
    cfg = wl.cfg
    if args.storage == "morton":
        from d2gs_b200 import layout
        layout.permute_surfels_(wl.pc, layout.morton_permutation(wl.pc.get_xyz))

    impl_note = None
    if args.impl == "ours":
        from d2gs_b200 import _lib, raster
        _lib.lib()   # fail loudly if the CUDA extension is missing
        raster.set_deferred_count(not args.sync_count)
        if args.tile_sort is not None:
            _lib.set_option("tile_sort", args.tile_sort)
        if wl.use_deform:
            build_deform_ours(wl)
        step_fn = step_ours
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import util as tutil
        wl.ref_mod = tutil.load_reference_ext()
        if wl.ref_mod is None:
            cb = cpu_baseline(args.config)
            print(json.dumps({"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 0, "steps": 1, "warmup": 0,
                              "ms_per_step": 1e3 / cb["value"], "higher_is_better": True, "impl": "reference",
                              "note": "oracle/_ref not loadable: timed the CPU oracle port instead", "cpu_baseline": dict(cb),
                              "config": {"workload": args.config},
                              "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return 0
        if wl.use_deform:
            build_deform_reference(wl)
        step_fn = step_reference
        wl.ref_timer = RefStageTimer(wl.ref_mod)
        wl.ref_timer.install()
        impl_note = "unmodified reference CUDA rasterizer (oracle/_ref) inside the reference's eager-torch pipeline (oracle/reference_pipeline.py)"

    from d2gs_b200 import dist as ddist
    params = list(wl.pc.raster_parameters()) + list(wl.deform_parameters())
    # flat gradient bucket: our backward kernels write parameter gradients straight into it (dist.claim), the reference arm's
    # autograd accumulates into pre-attached views; either way the all-reduce needs no pack step
    # (surfel tables whose gradient is final after the rasterizer backward start their all-reduce early, overlapped with
    # the deformation / MLP backward; `feature` also feeds the deformation blend, so it is not among them)
    early = [p for p in wl.pc.raster_parameters() if p is not getattr(wl.pc, "feature", None)]
    stages = {"raster": early}
    if args.impl == "ours" and wl.use_deform:
        # final after the deformation-blend backward (before the MLP backward): the hyper-coordinate table and the node geometry
        cn = wl.deform.deform
        stages["deform"] = [p for p in (wl.pc.feature, cn.nodes, cn._node_radius, cn._node_weight) if p.requires_grad]
    bucket = ddist.FlatGradBucket(params, direct=(args.impl != "reference"), stages=stages if args.early_allreduce else None)

    # N > 1: every step waits for its slowest rank, so the `world` views of a step are chosen to cost the same (instance
    # count of the view, measured once at set-up; dist.balanced_view_schedule).  Each view is still visited once per epoch.
    schedule = None
    if world > 1 and args.view_schedule == "balanced" and args.impl == "ours":     # (the reference arm is single-GPU: plain order)
        from gaussian_renderer import render as _render_ours
        costs = []
        with torch.no_grad():
            for c in wl.cams:
                d_ = wl.deform.step(wl.pc.get_xyz.detach(), wl.deform.deform.expand_time(c.fid), feature=wl.pc.feature,
                                    motion_mask=wl.pc.motion_mask) if wl.use_deform else {"d_xyz": 0.0, "d_rotation": 0.0, "d_scaling": 0.0}
                o_ = _render_ours(c, wl.pc, wl.pipe, wl.bg, d_["d_xyz"], d_["d_rotation"], d_["d_scaling"])
                costs.append(int((o_["radii"].long() ** 2).sum()))          # ~ screen area of the surfels ~ instance count
        schedule = ddist.balanced_view_schedule(costs, world)

    def view_index(step):
        if schedule is not None:
            return schedule[step % len(schedule)][rank]
        return ddist.view_for(step, rank, world, N_VIEWS)

    # ---- training tail (--train): what train_gui.py does between loss.backward() and the next iteration (:388-432)
    train_tail = None
    if args.train:
        pc = wl.pc
        P_ = pc._xyz.shape[0]
        # learning rates of arguments/__init__.py scaled by 1e-3: the optimiser does the same work per step, but 100+ steps
        # against a random target image must not dissolve the synthetic scene the metric is quoted on
        LR = 1e-3
        groups = [{"params": [pc._xyz], "lr": 1.6e-4 * 5 * LR, "name": "xyz"}, {"params": [pc._features_dc], "lr": 2.5e-3 * LR, "name": "f_dc"},
                  {"params": [pc._features_rest], "lr": 2.5e-3 / 20 * LR, "name": "f_rest"}, {"params": [pc._opacity], "lr": 0.05 * LR, "name": "opacity"},
                  {"params": [pc._scaling], "lr": 5e-3 * LR, "name": "scaling"}, {"params": [pc._rotation], "lr": 1e-3 * LR, "name": "rotation"},
                  {"params": [pc.feature], "lr": 1e-2 * LR, "name": "feature"}]
        dgroups = [{"params": list(wl.deform_parameters()), "lr": 1.6e-4 * 5 * LR, "name": "deform"}] if wl.use_deform else []
        if args.impl == "ours":
            from d2gs_b200.optim import FusedAdam, add_densification_stats
            opt_s = FusedAdam(groups, lr=0.0, eps=1e-15)
            opt_d = FusedAdam(dgroups, lr=0.0, eps=1e-15) if dgroups else None
        else:
            opt_s = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
            opt_d = torch.optim.Adam(dgroups, lr=0.0, eps=1e-15) if dgroups else None
        accum = torch.zeros((P_, 1), device=device)
        denom = torch.zeros((P_, 1), device=device)
        pc.max_radii2D = torch.zeros((P_,), dtype=torch.int32, device=device)

        def train_tail():
            out = wl.last_out
            vs, vis, radii = out["viewspace_points"], out["visibility_filter"], out["radii"]
            if world > 1:
                # replicas must take identical densification decisions: SUM of the per-view statistics, MAX of the radii
                gn = torch.linalg.vector_norm(vs.grad[:, :2], dim=-1)
                a_, c_, r_ = ddist.reduce_densification_stats(gn, vis, radii)
                accum.add_(a_.view(-1, 1)); denom.add_(c_.view(-1, 1))
                torch.maximum(pc.max_radii2D, r_.to(pc.max_radii2D.dtype), out=pc.max_radii2D)
            elif args.impl == "ours":
                add_densification_stats(accum, denom, vs, vis)
                torch.maximum(pc.max_radii2D, radii * vis, out=pc.max_radii2D)
            else:
                accum[vis] += torch.norm(vs.grad[vis, :2], dim=-1, keepdim=True)       # scene/gaussian_model.py:484-486
                denom[vis] += 1
                pc.max_radii2D[vis] = torch.max(pc.max_radii2D[vis], radii[vis])     # train_gui.py:389-391
            opt_s.step()
            if opt_d is not None:
                opt_d.step()

    # One device camera whose tensors are views of a single 36-float block (view 16 | proj 16 | centre 3 | time 1): a step
    # selects its view with ONE small copy into it — from the device-resident table (`value`) or from pinned host memory (`e2e`).
    def cam_block(c):
        return torch.cat([torch.as_tensor(c.world_view_transform, dtype=torch.float32).reshape(-1),
                          torch.as_tensor(c.full_proj_transform, dtype=torch.float32).reshape(-1),
                          torch.as_tensor(c.camera_center, dtype=torch.float32).reshape(-1), torch.tensor([c.fid], dtype=torch.float32)])
    host_blocks = [cam_block(c).pin_memory() for c in wl.cams_np]
    dev_blocks = torch.stack(host_blocks).to(device)
    step_block = torch.empty(36, dtype=torch.float32, device=device)
    step_cam = mdl.ViewCamera(wl.cams_np[0], device)
    step_cam.world_view_transform = step_block[0:16].view(4, 4)
    step_cam.full_proj_transform = step_block[16:32].view(4, 4)
    step_cam.camera_center = step_block[32:35]
    step_cam.fid = step_block[35:36]
    stage_block = torch.empty(36, dtype=torch.float32).pin_memory()     # e2e: the host writes the step's camera here
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()        # e2e: the step's result lands here
    e2e_gt = torch.empty_like(wl.gt_dev)
    copy_stream = torch.cuda.Stream(device)
    gt_ready = torch.cuda.Event()

    def step_body(e2e):
        """Device work of one step, issued on the current stream (eagerly, or once under CUDA-graph capture)."""
        bucket.zero()
        if e2e:
            # per-step inputs come from pinned host memory: camera block and the 7.7 MB target image; the image is only
            # needed by the loss, so it travels on a copy stream overlapped with the deformation and the forward render
            copy_stream.wait_stream(torch.cuda.current_stream())
            step_block.copy_(stage_block, non_blocking=True)
            with torch.cuda.stream(copy_stream):
                e2e_gt.copy_(wl.gt_host, non_blocking=True)
                gt_ready.record(copy_stream)
            loss = step_fn(wl, step_cam, e2e_gt, gt_ready)
        else:
            loss = step_fn(wl, step_cam, wl.gt_dev)
        bucket.all_reduce()      # finalize + (N > 1) ONE NCCL all-reduce over the flat buffer
        if e2e:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        return loss

    graphs = {}

    def run_step(step, e2e=False):
        vi = view_index(step)
        g = graphs.get(e2e)
        if e2e:
            stage_block.copy_(host_blocks[vi])           # host -> pinned staging (the previous step ended with a host read)
        else:
            step_block.copy_(dev_blocks[vi], non_blocking=True)
        if g is not None:
            g.replay()
        else:
            step_body(e2e)
        if train_tail is not None:
            train_tail()      # eager: the Adam step sizes are host-computed per step (bias correction), 2 launches
        if e2e:
            torch.cuda.current_stream().synchronize()    # device -> host read of the step's result
            return float(loss_host[0])
        return None

    def capture(e2e):
        """The step as ONE CUDA graph, recorded through the same public calls (DeformModel.step, render, loss, backward)."""
        torch.cuda.synchronize(device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            step_body(e2e)
        graphs[e2e] = g

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(n_steps, e2e, first_step):
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and not args.no_clocks:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for s in range(n_steps):
            run_step(first_step + s, e2e=e2e)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = (sampler.stop() if not args.no_clocks else {"sm_mhz": None, "reasons": ["not sampled"]}) if rank == 0 else None
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks

    for s in range(args.warmup):
        run_step(s)
    use_graph = args.impl == "ours" and args.graph != "off"
    graph_note = None
    if use_graph:
        try:
            capture(False)
            for s in range(3):
                run_step(s)
            torch.cuda.synchronize(device)
        except Exception as ex:      # noqa: BLE001 — capture problems must not hide the eager number
            if args.graph == "on":
                raise
            graphs.pop(False, None)
            graph_note = f"CUDA-graph capture failed, eager launches timed instead: {type(ex).__name__}: {str(ex)[:200]}"
            log(graph_note)
            torch.cuda.synchronize(device)
    graphed = graphs.get(False) is not None
    if args.impl == "ours" and not graphed:      # eager timed region: the stage timers run inside it
        from d2gs_b200 import _lib
        _lib.profile_collect()
        _lib.profile_enable(True)
    ms, clocks = timed(args.steps, False, args.warmup)
    stage, ms_eager = None, None
    if args.impl == "ours" and not graphed:
        stage, ms_eager = _lib.profile_collect(), ms
        _lib.profile_enable(False)
    value = args.steps * world / (ms / 1e3)

    # end to end: per-step inputs from pinned host memory, result read back by the host, every step
    for s in range(3):
        run_step(s, e2e=True)
    if use_graph and graphed:
        try:
            capture(True)
            for s in range(3):
                run_step(s, e2e=True)
        except Exception as ex:      # noqa: BLE001
            graphs.pop(True, None)
            graph_note = f"CUDA-graph capture of the e2e step failed, eager launches timed instead: {type(ex).__name__}: {str(ex)[:200]}"
            log(graph_note)
            torch.cuda.synchronize(device)
    ms_e2e, _ = timed(args.steps, True, args.warmup)
    e2e_value = args.steps * world / (ms_e2e / 1e3)
    e2e_graphed = graphs.get(True) is not None

    # per-stage CUDA-event timers (roofline) and launch counts: K eager steps of the same workload, stage events around
    # every launch.  (When the timed region replays a CUDA graph it contains the same launches, but events inside a graph
    # cannot be read per launch, so the stage timers run right after it.)
    ref_stages = None
    if args.impl == "ours":
        from d2gs_b200 import _lib, raster
        if graphed:
            graphs.clear()
            for s_ in range(3):
                run_step(s_)
            _lib.profile_collect()
            _lib.profile_enable(True)
            ms_eager, _ = timed(args.steps, False, args.warmup)
            stage = _lib.profile_collect()
            _lib.profile_enable(False)
        raster.check_deferred_counts(device)     # a frame that overflowed its binning capacity would have rendered NaN
        R = raster.last_num_rendered(device, cfg["P"], cfg["W"], cfg["H"])
    else:
        R = None
        # stage breakdown of the reference arm: the same K steps once more with CUDA events around the eager deformation,
        # the reference extension's two native calls and the step (not part of the timed region above)
        wl.ref_timer.on = True
        n_st = min(args.steps, 20)
        for s_ in range(n_st):
            with wl.ref_timer.mark("step"):
                run_step(args.warmup + s_)
        ref_stages = wl.ref_timer.collect(n_st)
        wl.ref_timer.on = False

    # rasterizer-only fwd+bwd on the same scene with the deformation off (the comparison north_star's ">= 3x" names)
    ronly = None
    if rank == 0 and not args.no_raster_only:
        try:
            torch.cuda.empty_cache()
            if args.impl == "ours":
                from d2gs_b200 import raster as _r
                _r.set_deferred_count(False)          # the op's default: what a trainer that only swaps the package gets
                mine = raster_only(args.config, device, "ours")
                ronly = {"ours_ms": mine["ms"], "ref_ms": None, "ratio": None, "workload": mine["workload"]}
                # baseline leg (like cpu_baseline): the unmodified reference extension, when it travelled to this box
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import util as tutil
                ref_mod = tutil.load_reference_ext()
                if ref_mod is not None:
                    theirs = raster_only(args.config, device, "reference", ref_mod=ref_mod)
                    ronly.update(ref_ms=theirs["ms"], ratio=theirs["ms"] / mine["ms"])
                else:
                    ronly["note"] = "oracle/_ref not loadable on this box: reference side not timed"
            else:
                theirs = raster_only(args.config, device, "reference", ref_mod=wl.ref_mod)
                ronly = {"ref_ms": theirs["ms"], "workload": theirs["workload"]}
        except Exception as ex:      # noqa: BLE001 — must not hide the headline number
            ronly = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
    h2d = 36 * 4 + 3 * wl.H * wl.W * 4

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    P, HW = cfg["P"], cfg["W"] * cfg["H"]
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (seeded D-NeRF-shaped scene, random-init deform MLP)",
           # `config` is the WORKLOAD and is identical in both arms; how each arm executes it is in `execution`
           "config": {"workload": f"{args.config}: {P} surfels + {cfg['n_nodes']} control nodes (K={cfg['K']}, hyper_dim 8, local_frame), "
                                  f"SH3, {cfg['W']}x{cfg['H']}, {N_VIEWS} views, 1 view/GPU/step, deform+render+loss+backward"
                                  + (" + densification stats + Adam step of all parameters" if args.train else ""),
                      "loss": "seeded random-weighted sum over the render outputs + L1 (SURVEY 8(d))" if args.loss == "synthetic" else
                              "training loss of train_gui.py:292-313: L1 + D-SSIM(0.2) + normal(0.02) + distortion(1000)",
                      "parallelism": f"view-sharded x{world}",
                      "storage": ("surfel tables Morton-sorted once at set-up (d2gs_b200.layout.permute_surfels_, both arms)"
                                  if args.storage == "morton" else "surfel tables in generation (random) order"),
                      "l2_policy": "no explicit flush: per-step working set (params+grads+workspaces ~0.3 GB) exceeds the 126 MB L2 and the view changes every step"},
           "execution": {"views": ("the views of one step (one per rank) are chosen to cost the same (dist.balanced_view_schedule); every view once per epoch"
                                   if schedule is not None else "view (step * N + rank) mod 100"),
                         "binning": ("reference extension: global radix sort, synchronous count readback" if args.impl == "reference" else
                                     ("global radix sort" if args.tile_sort == 0 else "per-tile buckets + per-tile sort") + ", " +
                                     ("synchronous count readback" if args.sync_count else "deferred count (no host synchronisation in the step)")),
                         "launch": ("one CUDA graph per step (captured through the public API)" if graphed else "eager launches") +
                                   (" | e2e: graph incl. H2D/D2H copies" if e2e_graphed else " | e2e: eager"),
                         "collective": ("NCCL all-reduce of the flat gradient bucket" + (", surfel tables after the rasterizer backward, hyper-coordinate table + node geometry after the deformation backward" if args.early_allreduce else "")) if world > 1 and args.impl == "ours"
                                       else ("rank 0 only (the reference is single-GPU)" if world > 1 else "none")},
           "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps}}
    if args.impl == "reference":
        out["impl"] = "reference"
        out["note"] = impl_note
        out["gpu_launches"] = 0
        out["stages_ms"] = ref_stages
        out["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                               "sample": "the reference has no CPU rasterizer (rasterize_points.cu:27-28 asserts CUDA); this arm runs the unmodified reference CUDA extension on the same B200"}
    else:
        tiles = ((cfg["W"] + 15) // 16) * ((cfg["H"] + 15) // 16)
        n_pass = math.ceil((32 + max(1, math.ceil(math.log2(tiles)) + 1)) / 8)
        sb = stage_bytes(P, R, HW, n_pass)
        mine = {k: v for k, v in stage.items() if k in sb and v[1] > 0}
        dom = max(mine, key=lambda k: mine[k][0])
        avg_ms = mine[dom][0] / mine[dom][1]
        achieved = sb[dom] / (avg_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom)
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                           "frac": achieved / hbm_peak, "traffic": traffic,
                           "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture (profiles/ncu_r2.md), not measured in this run",
                           "avg_launch_ms": avg_ms,
                           "algorithmic_bytes_per_launch": sb[dom], "peak_source": peak_src,
                           "frame": {"algorithmic_bytes": frame_bytes(P, R, HW, wl.use_deform),
                                     "achieved": frame_bytes(P, R, HW, wl.use_deform) / (ms / args.steps * 1e-3) / 1e9,
                                     "frac": frame_bytes(P, R, HW, wl.use_deform) / (ms / args.steps * 1e-3) / 1e9 / hbm_peak},
                           "stages_ms": {k: (v[0] / v[1] if v[1] else None) for k, v in stage.items()}}
        out["num_rendered"] = R
        out["eager_ms_per_step"] = ms_eager / args.steps
        if graph_note:
            out["graph_note"] = graph_note
        out["roofline"]["timing"] = ("CUDA events around every launch of the stage, averaged over %d eager steps run right after the "
                                     "timed region (which replays the same launches as a CUDA graph)" % args.steps) if graphed else \
                                    "CUDA events around every launch of the stage inside the timed region"
        # hand-written kernels launched per stage call (CUB scan/sort launches are library code and not counted)
        mine_kernels = {"preprocess_fwd": 1, "duplicate": 1, "ranges": 1, "blend_fwd": 1, "blend_bwd": 1, "preprocess_bwd": 1,
                        "deform_fwd": 1, "deform_bwd": 1, "epilogue_fwd": 1, "epilogue_bwd": 1, "mlp_fwd": 2, "mlp_bwd": 2,
                        "loss_fwd": 2, "loss_bwd": 1}
        if args.tile_sort != 0:   # tile-bucketed binning (library default): count + scan | scatter | per-tile sort (2 kernels), no CUB
            mine_kernels.update({"scan": 2, "duplicate": 1, "sort": 2})
        out["gpu_launches"] = int(sum(stage[k][1] * n for k, n in mine_kernels.items() if k in stage))
        if args.train:      # + FusedAdam (one launch per optimiser) and the densification-statistics kernel (single GPU)
            out["gpu_launches"] += args.steps * ((2 if wl.use_deform else 1) + (1 if world == 1 else 0))
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(args.config)
            except Exception as ex:   # the oracle is a checker; its absence must not hide the GPU number
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if ronly is not None:
        out["raster_only"] = ronly
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
