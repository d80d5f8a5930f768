#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/n2c.log 2>&1
date
run() {  # name, extra env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 8 --no-cpu-baseline --no-raster-only > gpurun_out/n2c_$name.json 2> gpurun_out/n2c_$name.err
  python - <<PY
import json
txt=open("gpurun_out/n2c_$name.json").read()
try:
    d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1]); s=d["roofline"]["stages_ms"]
    print("$name", "ms/step %.4f"%d["ms_per_step"], "frames/s %.1f"%d["value"], "deform_bwd %.4f mlp_bwd %.4f mlp_fwd %.4f"%(s["deform_bwd"], s["mlp_bwd"], s["mlp_fwd"]))
except Exception as e:
    print("$name", "ERR", e)
PY
}
run default A=1
run bwdcl4 D2GS_OPTIONS=mlp_cluster_bwd=4
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run ctas4 NCCL_MAX_CTAS=4
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
date
