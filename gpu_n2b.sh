#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/n2b.log 2>&1
date
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_gs3d_gpu.py -m gpu -q 2>&1 | grep -v "^  \|^$" | cut -c1-400 | tail -12
date
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 8 --no-cpu-baseline --no-raster-only > gpurun_out/n2b_ours.json 2> gpurun_out/n2b.err; tail -3 gpurun_out/n2b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 8 --no-cpu-baseline --no-raster-only --early-allreduce 0 > gpurun_out/n2b_ours_noearly.json 2> gpurun_out/n2b_noearly.err; tail -3 gpurun_out/n2b_noearly.err
timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-raster-only > gpurun_out/n2b_ours_n1.json 2> gpurun_out/n2b_n1.err
date
python - <<'PY'
import json
for f in ("n2b_ours_n1", "n2b_ours", "n2b_ours_noearly"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); print(f, d["n_gpus"], "gpus", "ms/step %.4f" % d["ms_per_step"], "frames/s %.1f" % d["value"], d["execution"].get("collective"))
    except Exception as e:
        print(f, "ERR", e)
PY
