#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/n8c.log 2>&1
date
run() {  # name, bench args...
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline --no-raster-only "$@" > gpurun_out/n8c_$name.json 2> gpurun_out/n8c_$name.err
  python - <<PY
import json
txt=open("gpurun_out/n8c_$name.json").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); s=d["roofline"]["stages_ms"]
    print("$name", d["n_gpus"], "gpus ms/step %.4f"%d["ms_per_step"], "frames/s %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], d["config"]["parallelism"])
except Exception as e:
    print("$name", "ERR", e); print(open("gpurun_out/n8c_$name.err").read()[-1500:])
PY
}
run balanced --steps 52 --warmup 13
run roundrobin --steps 52 --warmup 13 --view-schedule roundrobin
date
