#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/pytest_k3.log 2>&1; tail -3 gpurun_out/pytest_k3.log
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -3 gpurun_out/pytest_net3.log
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2g_shape_pong_tc3.txt 2>&1; head -14 gpurun_out/r2g_shape_pong_tc3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_kernel -s 2 -c 1 -o gpurun_out/r2g_tc3_conv2 -f python scratch/one_conv.py tc3 0 8192 20 20 32 64 4 4 2 0 > gpurun_out/ncu_conv2.log 2>&1; tail -3 gpurun_out/ncu_conv2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_kernel -s 2 -c 1 -o gpurun_out/r2g_tc3_conv1 -f python scratch/one_conv.py tc3 0 8192 21 21 64 64 2 2 1 0 > gpurun_out/ncu_conv1.log 2>&1; tail -3 gpurun_out/ncu_conv1.log
ls -la gpurun_out/*.ncu-rep
