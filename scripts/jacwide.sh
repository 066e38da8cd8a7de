# Wide-block (16x16 pivot) Jacobi: accuracy on odd sizes, timing at m = 513/1025/2049 (NLS_JACOBI_JB=4: old variant).
mkdir -p gpurun_out
{
echo "== JB=8 small"; NLS_JACOBI_JB=8 timeout 300 python tests/gpu_diag.py eig_small 2>&1 | grep "eig_small"
echo "== JB=8"; NLS_JACOBI_JB=8 timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi"
echo "== JB=8 diag=1 (no pivot solves, 10 sweeps)"; NLS_JACOBI_DIAG=1 timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi" | cut -c1-60
echo "== JB=8 diag=2 (no tile updates, 10 sweeps)"; NLS_JACOBI_DIAG=2 timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi" | cut -c1-60
echo "== JB=8 diag=3 (empty rounds, 10 sweeps)"; NLS_JACOBI_DIAG=3 timeout 300 python tests/gpu_diag.py eig 2>&1 | grep "eig/jacobi" | cut -c1-60
} | tee gpurun_out/jacwide.log
