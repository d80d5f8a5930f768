"""Diagnosis (not a pytest): where does a bench step spend its time?  Host wall-clock per section with syncs."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-2dgs_b200"))
import torch
import bench
from gaussian_renderer import render

dev = torch.device("cuda:0")
wl = bench.Workload("C3", dev, "ours")
bench.build_deform_ours(wl)
params = list(wl.pc.raster_parameters()) + list(wl.deform_parameters())
def sync(): torch.cuda.synchronize()
def section(acc, name, t0):
    sync(); t = time.perf_counter(); acc[name] = acc.get(name, 0) + (t - t0); return t
for mode in ("sync-sections", "free-running"):
    acc = {}
    n = 20
    for it in range(n + 5):
        if it == 5:
            acc = {}; sync(); T0 = time.perf_counter()
        for p in params: p.grad = None
        cam = wl.cams[it % 100]
        pc = wl.pc
        if mode == "sync-sections":
            sync(); t = time.perf_counter()
            t_in = wl.deform.deform.expand_time(cam.fid)
            d = wl.deform.step(pc.get_xyz.detach(), t_in, feature=pc.feature, motion_mask=pc.motion_mask)
            t = section(acc, "deform_fwd", t)
            out = render(cam, pc, wl.pipe, wl.bg, d["d_xyz"], d["d_rotation"], d["d_scaling"])
            t = section(acc, "render_fwd", t)
            loss = bench.synthetic_loss(out, wl.wts, wl.gt_dev)
            t = section(acc, "loss_fwd", t)
            loss.backward()
            t = section(acc, "backward", t)
        else:
            bench.step_ours(wl, cam, wl.gt_dev)
    sync(); T1 = time.perf_counter()
    print(mode, "ms/step", 1e3 * (T1 - T0) / n, {k: round(1e3 * v / n, 3) for k, v in acc.items()})

# host-only cost: count launches via profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for it in range(3):
        for p in params: p.grad = None
        bench.step_ours(wl, wl.cams[it], wl.gt_dev)
    sync()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=60))
