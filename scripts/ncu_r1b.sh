mkdir -p gpurun_out
# full capture: Gram + feature map (transposed) of pass 1, Jacobi kernels, bin-statistics kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"OpGram|jacobi_pivot|jacobi_update|bs_count|bs_stats_kernel|bs_mad" -s 2 -c 12 -o gpurun_out/prof_misc_r1 python scripts/profile_fit_api.py 300000 > gpurun_out/ncu_misc.log 2>&1
tail -3 gpurun_out/ncu_misc.log | cut -c1-200
