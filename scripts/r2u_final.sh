#!/bin/bash
# round 2 final evidence on one B200: full GPU suite, smoke, bench (N = 1, full line), reference arm, ncu launch list,
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2u_launches_bench_262k.csv python bench.py --steps 1 --warmup 1 --rows 262144 --skip-api --skip-configs > gpurun_out/r2u_ncu_bench.log 2>&1
tail -3 gpurun_out/r2u_pytest.log; tail -1 gpurun_out/r2u_smoke.log; cut -c1-260 gpurun_out/r2u_bench_n1.json; cut -c1-200 gpurun_out/r2u_bench_reference.json; wc -l gpurun_out/r2u_launches_bench_262k.csv
