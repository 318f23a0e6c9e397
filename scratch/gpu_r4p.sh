#!/bin/bash
# 4 GPUs with the final kernels (own all-reduce kernel)
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 4 --steps 6 --warmup 3 --no-cpu --no-others > gpurun_out/r4p_bench_dp4.json 2> gpurun_out/r4p_bench_dp4.err; head -c 300 gpurun_out/r4p_bench_dp4.json; echo; tail -n 2 gpurun_out/r4p_bench_dp4.err
