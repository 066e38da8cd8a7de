# Round-1 (g): final validation — GPU tests, smoke(), N=1 bench with the public-API arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_err.log; tail -c 1500 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_err.log
