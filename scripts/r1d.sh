# Round-1 (d): GPU tests, N=1 bench (h2d-overlapped e2e arm), ncu launch list of a reduced-row bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_err.log; tail -c 2600 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 6000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --rows 262144 --steps 1 --warmup 1 --skip-api > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r1d.csv
