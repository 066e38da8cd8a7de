#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiProject|OpSweep|slice_rows" -c 6 -o gpurun_out/r2g_ncu_full_project_sweep -f python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2g_ncu_full.log 2>&1
tail -3 gpurun_out/r2g_ncu_full.log
python -m pytest tests/test_gpu_estimator.py -m gpu -q -k "nan or large" 2>&1 | tail -3
