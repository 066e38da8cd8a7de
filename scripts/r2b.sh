#!/bin/bash
# round 2b: first run of the tridiagonalisation + divide-and-conquer eigensolver on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_primal.py -q -x -k "dc_eigensolver or heev_matches" 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
timeout 900 python scripts/eig_bench.py --complex 513,1025,2049,4097 --real 4096 > gpurun_out/r2b_eig_bench.log 2>&1
tail -5 gpurun_out/r2b_pytest.log; tail -20 gpurun_out/r2b_eig_bench.log
