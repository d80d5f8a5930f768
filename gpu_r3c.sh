#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r3c.log 2>&1
date
V=dynamic-2dgs_b200/build/variants
timeout 900 python tests/gpu_ab.py --cfg C3 default $V/libd2gs_fw8.so $V/libd2gs_bw8r64.so $V/libd2gs_bw8r80.so default
date
