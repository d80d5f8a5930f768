#!/bin/bash
# 8-GPU configurations of DESIGN.md section 7 (gpurun --gpus 8).
mkdir -p gpurun_out
exec > gpurun_out/n8b.log 2>&1
date
nvidia-smi -L | wc -l
run() {  # name, bench args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline --no-raster-only "$@" > gpurun_out/n8b_$name.json 2> gpurun_out/n8b_$name.err
  python - <<PY
import json
txt=open("gpurun_out/n8b_$name.json").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); s=d["roofline"]["stages_ms"]
    print("$name", d["n_gpus"], "gpus ms/step %.4f"%d["ms_per_step"], "frames/s %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], "deform_bwd %.4f mlp_bwd %.4f"%(s["deform_bwd"] or 0, s["mlp_bwd"] or 0))
except Exception as e:
    print("$name", "ERR", e)
PY
}
run C3 --steps 40 --warmup 8
run C3_noearly --steps 40 --warmup 8 --early-allreduce 0
run C3_train --steps 40 --warmup 8 --train
run C5 --config C5 --steps 20 --warmup 6
date
