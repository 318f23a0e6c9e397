#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scratch/symm_probe.py > gpurun_out/r2z_symm_probe.log 2>&1; tail -n 3 gpurun_out/r2z_symm_probe.log
timeout 900 python -m pytest tests/test_gpu_dist.py -v -q > gpurun_out/r2z_pytest_dist_2gpu.log 2>&1; tail -n 25 gpurun_out/r2z_pytest_dist_2gpu.log
for c in peer nccl; do
DDRL_DP_COLLECTIVE=$c timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 --no-others > gpurun_out/r2z_bench_dp2_$c.json 2> gpurun_out/r2z_bench_dp2_$c.err; head -c 300 gpurun_out/r2z_bench_dp2_$c.json; echo; tail -n 3 gpurun_out/r2z_bench_dp2_$c.err
done
DDRL_DP_COLLECTIVE=peer timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 6 --warmup 3 --no-others > gpurun_out/r2z_bench_dp2_peer_b.json 2>/dev/null; head -c 300 gpurun_out/r2z_bench_dp2_peer_b.json; echo
