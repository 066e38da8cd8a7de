# Round-1 (e): final N=1 bench line + ncu capture of the Jacobi round kernel (two-warp pivots, staged tiles).
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_err.log; tail -c 3300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_err.log
timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:"jacobi_round_w" -s 300 -c 3 -o gpurun_out/prof_jacobi_wide_r1e python bench.py --rows 131072 --steps 1 --warmup 1 --skip-api > gpurun_out/ncu_jacw.log 2>&1
tail -2 gpurun_out/ncu_jacw.log
