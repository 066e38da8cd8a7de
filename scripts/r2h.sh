#!/bin/bash
# TMEM-A INT8 core in the library: GPU suite + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
cut -c1-250 gpurun_out/r2h_bench_n1.json; tail -3 gpurun_out/r2h_bench_n1.err
