mkdir -p gpurun_out
for occ in 4 5; do for inner in 1 2; do echo "== fused occ=$occ inner=$inner"; NLS_JACOBI_INNER=$inner NLS_JACOBI_OCC=$occ NLS_JACOBI_MODE=fused timeout 200 python tests/gpu_diag.py eig 2>&1 | grep -E "jacobi/m(513|1025)|EXC|rror" | cut -c1-150; done; done | tee gpurun_out/eig7.log
echo "== split"; NLS_JACOBI_MODE=split timeout 200 python tests/gpu_diag.py eig 2>&1 | grep -E "jacobi/m(513|1025)|EXC|rror" | cut -c1-100
timeout 600 python -m pytest tests/test_gpu_primal.py -x -q -k "heev or eigensolver or c1 or ragged" 2>&1 | tail -3
