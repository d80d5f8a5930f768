#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/n8d.log 2>&1
date
run() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/gpu_allreduce_probe.py 2>&1 | grep "ALLREDUCE_PROBE\|NVLS\|Error\|error" | cut -c1-700 | head -12; }
run NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
run NCCL_ALGO=Ring
run NCCL_ALGO=NVLS
run NCCL_NVLS_ENABLE=0
date
