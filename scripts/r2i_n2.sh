#!/bin/bash
# two GPUs: single-process sharding behind the estimator + torchrun bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2i_pytest_multi.log; cat gpurun_out/r2i_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/gpu_dist_check.py > gpurun_out/r2i_dist_check_n2.log 2>&1; tail -4 gpurun_out/r2i_dist_check_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
cut -c1-330 gpurun_out/r2i_bench_n2.json; tail -2 gpurun_out/r2i_bench_n2.err
