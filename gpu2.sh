mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_raster_gpu.py tests/test_render_gpu.py -q -m gpu --tb=short 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --steps 60 --warmup 8 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']); print(d['roofline']['stages_ms'])"
