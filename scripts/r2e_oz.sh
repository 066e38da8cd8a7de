#!/bin/bash
# INT8 projection core inside the library: the GPU tests that exercise it, then the bench with both cores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_primal.py tests/test_gpu_estimator.py tests/test_gpu_configs.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
python bench.py --steps 2 --warmup 3 --skip-api --skip-configs > gpurun_out/r2e_bench_ozaki.json 2> gpurun_out/r2e_bench_ozaki.err
NLS_GEMM=dmma python bench.py --steps 2 --warmup 3 --skip-api --skip-configs > gpurun_out/r2e_bench_dmma.json 2> gpurun_out/r2e_bench_dmma.err
cut -c1-1500 gpurun_out/r2e_bench_ozaki.json; tail -3 gpurun_out/r2e_bench_ozaki.err
cut -c1-300 gpurun_out/r2e_bench_dmma.json
