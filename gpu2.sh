mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -15 | cut -c1-250
python bench.py --steps 60 --warmup 8 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-200
