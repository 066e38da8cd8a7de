#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_primal.py tests/test_gpu_configs.py tests/test_gpu_estimator.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2j_pytest.log
tail -6 gpurun_out/r2j_pytest.log
python bench.py --steps 2 --warmup 2 --skip-api --skip-configs > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
cut -c1-250 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
