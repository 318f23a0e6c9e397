#!/bin/bash
# N GPUs (N = $1): peer all-reduce kernel check + timings, bench with the own collective and with NCCL
N=${1:-4}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -n 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 scratch/ar_check.py > gpurun_out/r3d_ar_check_n$N.log 2>&1; tail -n 9 gpurun_out/r3d_ar_check_n$N.log | cut -c1-300
for c in peer nccl; do
DDRL_DP_COLLECTIVE=$c timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $N --steps 6 --warmup 3 --no-others > gpurun_out/r3d_bench_dp${N}_$c.json 2> gpurun_out/r3d_bench_dp${N}_$c.err; head -c 300 gpurun_out/r3d_bench_dp${N}_$c.json; echo; tail -n 3 gpurun_out/r3d_bench_dp${N}_$c.err
done
