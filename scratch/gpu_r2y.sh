#!/bin/bash
set -x
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scratch/symm_probe.py > gpurun_out/r2y_symm_probe.log 2>&1; grep -v "NCCL INFO" gpurun_out/r2y_symm_probe.log | tail -n 15; grep -i "nvls" gpurun_out/r2y_symm_probe.log | head -n 5
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -k "segments or graph_replay" > gpurun_out/r2y_pytest_seg.log 2>&1; tail -n 5 gpurun_out/r2y_pytest_seg.log
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > gpurun_out/r2y_pytest_dist.log 2>&1; tail -n 5 gpurun_out/r2y_pytest_dist.log
