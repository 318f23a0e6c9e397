#!/bin/bash
# 2-GPU pass: segmented backward (single-GPU test), data-parallel parity tests with and without the all-reduce overlap,
# 2-GPU bench lines with the overlap on / off
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -k "segments or graph_replay or full_size" > gpurun_out/r2x_pytest_seg.log 2>&1; tail -n 5 gpurun_out/r2x_pytest_seg.log
timeout 900 python -m pytest tests/test_gpu_dist.py -v -x -q > gpurun_out/r2x_pytest_dist_2gpu.log 2>&1; tail -n 12 gpurun_out/r2x_pytest_dist_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 --no-others > gpurun_out/r2x_bench_dp2_overlap.json 2> gpurun_out/r2x_bench_dp2_overlap.err; head -c 600 gpurun_out/r2x_bench_dp2_overlap.json; echo; tail -n 5 gpurun_out/r2x_bench_dp2_overlap.err
DDRL_DP_OVERLAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 6 --warmup 3 --no-others > gpurun_out/r2x_bench_dp2_serial.json 2> gpurun_out/r2x_bench_dp2_serial.err; head -c 600 gpurun_out/r2x_bench_dp2_serial.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 6 --warmup 3 --no-others --workload navlaser > gpurun_out/r2x_bench_dp2_navlaser_overlap.json 2>/dev/null; head -c 400 gpurun_out/r2x_bench_dp2_navlaser_overlap.json; echo
DDRL_DP_OVERLAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 6 --warmup 3 --no-others --workload navlaser > gpurun_out/r2x_bench_dp2_navlaser_serial.json 2>/dev/null; head -c 400 gpurun_out/r2x_bench_dp2_navlaser_serial.json; echo
