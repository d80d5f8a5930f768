mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 2 -c 2 -f -o gpurun_out/prof_blend_r1g python tests/gpu_one_frame.py C3 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
