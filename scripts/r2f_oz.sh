#!/bin/bash
# INT8 core everywhere it applies (Gram, projection, predict_std): probe, full GPU suite, full bench line, ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/micro/ozaki_pipe.bin > gpurun_out/r2f_ozaki_pipe.log 2>&1; tail -3 gpurun_out/r2f_ozaki_pipe.log
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
cut -c1-300 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches_bench_262k.csv python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel_i8|slice_rows_kernel|slice_gram_kernel" -s 6 -c 8 -o gpurun_out/r2f_ncu_full_int8 -f python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2f_ncu_full.log 2>&1
ls -la gpurun_out/r2f_*
