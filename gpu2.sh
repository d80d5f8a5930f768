mkdir -p gpurun_out
export D2GS_TEST_REPORT=gpurun_out/pipeline_grad_report.json
timeout 900 python -m pytest tests/test_deform_gpu.py -q -m gpu --tb=short -x 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -15 | cut -c1-600
timeout 1200 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_deform_gpu.py 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --steps 60 --warmup 8 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
