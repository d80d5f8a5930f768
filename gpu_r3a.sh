#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r3a.log 2>&1
date
V=dynamic-2dgs_b200/build/variants
timeout 900 python tests/gpu_ab.py --cfg C3 default $V/libd2gs_fw2.so $V/libd2gs_bw2.so $V/libd2gs_fb64.so default
date
