"""Probe (not a pytest): time of the bucket-sized NCCL all-reduces on this box, eager and inside a CUDA graph.
    torchrun --nproc-per-node N tests/gpu_allreduce_probe.py"""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
sizes_mb = [2.1, 13.0, 59.0, 79.0]
res = {}
for mb in sizes_mb:
    n = int(mb * 1e6 / 4)
    x = torch.ones(n, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 20
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            dist.all_reduce(x)
        g.replay(); torch.cuda.synchronize(); dist.barrier()
        e0.record(s)
        for _ in range(20):
            g.replay()
        e1.record(s); torch.cuda.synchronize()
    graphed = e0.elapsed_time(e1) / 20
    res[mb] = (eager, graphed)
if rank == 0:
    print("ALLREDUCE_PROBE world=%d env=%s " % (world, {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and k != "NCCL_DEBUG"}) +
          " ".join(f"{mb}MB: eager {a:.3f} ms ({mb / a:.0f} GB/s) graph {b:.3f} ms" for mb, (a, b) in res.items()), flush=True)
dist.destroy_process_group()
