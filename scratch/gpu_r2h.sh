#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/pytest_k3.log 2>&1; tail -n 3 gpurun_out/pytest_k3.log
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -n 3 gpurun_out/pytest_net3.log
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2h_shape_pong_tc3.txt 2>&1; head -n 22 gpurun_out/r2h_shape_pong_tc3.txt
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing.so timeout 300 python scratch/tc3_roles.py > gpurun_out/r2h_tc3_roles.txt 2>&1; head -n 14 gpurun_out/r2h_tc3_roles.txt
