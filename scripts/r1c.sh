# Round-1 (c): GPU tests, N=1 bench with the wide Jacobi eigensolver, ncu capture of the Jacobi round kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_err.log; tail -c 3000 gpurun_out/bench_n1.json
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:"jacobi_round_w" -s 300 -c 3 -o gpurun_out/prof_jacobi_wide_r1 python bench.py --rows 131072 --steps 1 --warmup 1 --skip-api > gpurun_out/ncu_jacw.log 2>&1
tail -3 gpurun_out/ncu_jacw.log
ls -la gpurun_out/*.ncu-rep
