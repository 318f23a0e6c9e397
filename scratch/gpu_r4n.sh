#!/bin/bash
# 2 GPUs with the final kernels: dist parity tests + bench line (own all-reduce kernel)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/r4n_pytest_dist_2gpu.log 2>&1; tail -n 4 gpurun_out/r4n_pytest_dist_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/r4n_bench_dp2.json 2> gpurun_out/r4n_bench_dp2.err; head -c 300 gpurun_out/r4n_bench_dp2.json; echo; tail -n 3 gpurun_out/r4n_bench_dp2.err
