#!/bin/bash
# round 2 final evidence on one B200: full GPU suite, smoke, bench (N = 1, full line), reference arm, ncu launch list,
# ncu --set full of the INT8 GEMM kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2q_launches_bench_262k.csv python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2q_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiProjectPlanes|EpiSweep|EpiGram" -s 3 -c 6 -o gpurun_out/r2q_ncu_full_int8 -f python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2q_ncu_full.log 2>&1
tail -3 gpurun_out/r2q_pytest.log; tail -1 gpurun_out/r2q_smoke.log; cut -c1-260 gpurun_out/r2q_bench_n1.json; cut -c1-200 gpurun_out/r2q_bench_reference.json; wc -l gpurun_out/r2q_launches_bench_262k.csv; ls -la gpurun_out/r2q_ncu_full_int8.ncu-rep
