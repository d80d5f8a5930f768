mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_render_gpu.py tests/test_deform_gpu.py::test_render_dropin_matches_reference_pipeline -q -m gpu -x --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|DESIRED|ACTUAL" > gpurun_out/fail1.log 2>&1
timeout 600 python -m pytest tests/test_render_gpu.py::test_fused_epilogue_equals_eager tests/test_deform_gpu.py::test_render_dropin_matches_reference_pipeline -q -m gpu --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]" > gpurun_out/fail2.log 2>&1
python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_ours.json | head -2;  grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_ours.json
