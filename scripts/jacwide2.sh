mkdir -p gpurun_out
{
timeout 60 scripts/micro/pivot_bench.bin
for dg in 0 4; do echo "== layout diag=$dg"; NLS_JACOBI_DIAG=$dg timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi" | cut -c1-110; done
echo "== diag=2 (pivot path only)"; NLS_JACOBI_DIAG=2 timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi" | cut -c1-60
} | tee gpurun_out/jacwide2.log
timeout 600 python -m pytest tests/test_gpu_primal.py -m gpu -x -q -k "heev or jacobi or eigensolver" 2>&1 | tail -3
