mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -6 | cut -c1-600
timeout 600 python bench.py --steps 60 --warmup 8 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']); print(d['roofline']['stages_ms'])"
