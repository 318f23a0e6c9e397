#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/pytest_k3.log 2>&1; tail -15 gpurun_out/pytest_k3.log
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -25 gpurun_out/pytest_net3.log
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2c_shape_pong_tc3.txt 2>&1; head -45 gpurun_out/r2c_shape_pong_tc3.txt
DDRL_GEMM_MODE=tc2 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2c_shape_pong_tc2.txt 2>&1; head -30 gpurun_out/r2c_shape_pong_tc2.txt
timeout 600 python -m pytest tests/test_threads.py tests/test_gpu_encoders.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
