#!/bin/bash
# eight GPUs: the bench at N = 8, then C5 of BASELINE.json at full size (n = 16M, d = 128, m = 4097)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err
cut -c1-330 gpurun_out/r2n_bench_n8.json; tail -2 gpurun_out/r2n_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/gpu_c5_full.py > gpurun_out/r2n_c5_full.log 2>&1
tail -25 gpurun_out/r2n_c5_full.log; cp gpurun_out/c5_full.json gpurun_out/r2n_c5_full_n8.json 2>/dev/null
