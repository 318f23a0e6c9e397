#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -v -x -q > gpurun_out/r2k_pytest_dist_2gpu.log 2>&1; tail -n 8 gpurun_out/r2k_pytest_dist_2gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2k_bench_pong_dp2.json 2> gpurun_out/r2k_bench_pong_dp2.err; tail -c 3000 gpurun_out/r2k_bench_pong_dp2.json; tail -n 5 gpurun_out/r2k_bench_pong_dp2.err
