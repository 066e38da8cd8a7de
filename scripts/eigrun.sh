mkdir -p gpurun_out
for inner in 2 4 12; do echo "== inner=$inner"; NLS_JACOBI_INNER=$inner timeout 300 python tests/gpu_diag.py eig 2>&1 | grep jacobi; done | tee gpurun_out/eig.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
