mkdir -p gpurun_out
for pdl in 1 0; do echo "== pdl=$pdl"; NLS_JACOBI_PDL=$pdl timeout 200 python tests/gpu_diag.py eig 2>&1 | grep -E "jacobi/m(513|1025)|EXC|rror" | cut -c1-120; done | tee gpurun_out/eig3.log
timeout 600 python -m pytest tests/test_gpu_primal.py -x -q -k "heev or eigensolver or c1 or ragged" 2>&1 | tail -5
