#!/bin/bash
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r3k_bench_dp$N.json 2> gpurun_out/r3k_bench_dp$N.err; head -c 300 gpurun_out/r3k_bench_dp$N.json; echo; tail -n 3 gpurun_out/r3k_bench_dp$N.err
DDRL_DP_COLLECTIVE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 6 --warmup 3 --no-others > gpurun_out/r3k_bench_dp${N}_nccl.json 2> gpurun_out/r3k_bench_dp${N}_nccl.err; head -c 300 gpurun_out/r3k_bench_dp${N}_nccl.json; echo
