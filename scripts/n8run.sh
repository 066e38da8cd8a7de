# 8-GPU evidence: NCCL parity check + C3 bench (strong scaling), one process per GPU.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_dist_check.py 2>&1 | grep "\[dist\]" | tee gpurun_out/dist_check_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 2500 gpurun_out/bench_n8.json
