mkdir -p gpurun_out
nvidia-smi -L | wc -l
for e in 0 1; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$e bench.py --gpus 4 --steps 40 --warmup 8 --early-allreduce $e --no-cpu-baseline > gpurun_out/bench_n4_e$e.json 2> gpurun_out/bench_n4_e$e.err; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": 4[^}]*ms_per_step": [0-9.]*' gpurun_out/bench_n4_e$e.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n4_e$e.json; tail -2 gpurun_out/bench_n4_e$e.err | cut -c1-200
done
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -iE "NVLS|nChannels|Connected all|Ring|Tree" | head -12 | cut -c1-200
