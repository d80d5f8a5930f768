mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -vE "^E\s+\[|^E\s+[0-9\.\-e, ]+\]|^E\s+\+|array\(" | tail -12 | cut -c1-600
for L in synthetic train; do
timeout 600 python bench.py --steps 60 --warmup 8 --loss $L --no-cpu-baseline > gpurun_out/bench_ours_$L.json 2> gpurun_out/bench_ours_$L.err; tail -3 gpurun_out/bench_ours_$L.err; cat gpurun_out/bench_ours_$L.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']); print(d['roofline']['stages_ms'])"
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 4 --loss train > gpurun_out/bench_ref_train.json 2> gpurun_out/bench_ref_train.err; tail -3 gpurun_out/bench_ref_train.err; cat gpurun_out/bench_ref_train.json | cut -c1-300
