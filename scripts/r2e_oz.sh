#!/bin/bash
# INT8 (Ozaki) Gram + projection inside the library: the full GPU suite, then the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
python bench.py --steps 2 --warmup 3 --skip-api --skip-configs > gpurun_out/r2e_bench_ozaki.json 2> gpurun_out/r2e_bench_ozaki.err
cut -c1-400 gpurun_out/r2e_bench_ozaki.json; tail -3 gpurun_out/r2e_bench_ozaki.err
