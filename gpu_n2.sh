mkdir -p gpurun_out
for e in 0 1 0 1; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$e bench.py --gpus 2 --steps 40 --warmup 8 --early-allreduce $e --no-cpu-baseline > gpurun_out/bench_n2_e$e.json 2> gpurun_out/bench_n2_e$e.err; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": 2[^}]*ms_per_step": [0-9.]*' gpurun_out/bench_n2_e$e.json; grep -o '"deform_bwd": [0-9.]*' gpurun_out/bench_n2_e$e.json
done
