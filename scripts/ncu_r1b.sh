mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"OpGram" -s 1 -c 2 -o gpurun_out/prof_gram_r1 python bench.py --rows 131072 --steps 1 --warmup 1 --skip-api > gpurun_out/ncu_gram.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"jacobi_pivot|jacobi_update" -s 600 -c 4 -o gpurun_out/prof_jacobi_r1 python bench.py --rows 131072 --steps 1 --warmup 1 --skip-api > gpurun_out/ncu_jac.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bs_count|bs_stats_kernel|bs_mad" -s 70 -c 6 -o gpurun_out/prof_bs_r1 python scripts/profile_fit_api.py 2000000 > gpurun_out/ncu_bs.log 2>&1
ls -la gpurun_out/*.ncu-rep
