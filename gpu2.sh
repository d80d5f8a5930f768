set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend_ -s 2 -c 2 -o gpurun_out/prof_blend_r1a python tests/gpu_one_frame.py C3 2 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
