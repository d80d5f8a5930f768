"""Scratch profiling target (not a pytest): host-side cost of one bench step, cProfile over N steps of the C3 workload."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dynamic-2dgs_b200"))
import torch
import bench
from d2gs_b200 import dist as ddist
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
wl = bench.Workload("C3", dev, "ours"); wl.loss_kind = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
bench.build_deform_ours(wl)
params = list(wl.pc.raster_parameters()) + list(wl.deform_parameters())
bucket = ddist.FlatGradBucket(params)
def step(i):
    bucket.zero()
    loss = bench.step_ours(wl, wl.cams[i % 100], wl.gt_dev)
    bucket.all_reduce()
for i in range(10): step(i)
torch.cuda.synchronize()
N = 100
t0 = time.perf_counter(); c0 = time.process_time()
for i in range(N): step(i)
torch.cuda.synchronize()
print(f"wall {1e3*(time.perf_counter()-t0)/N:.3f} ms/step, process cpu {1e3*(time.process_time()-c0)/N:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for i in range(N): step(i)
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
