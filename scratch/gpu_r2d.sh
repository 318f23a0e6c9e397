#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing.so timeout 300 python scratch/tc3_roles.py > gpurun_out/r2d_tc3_roles.txt 2>&1; cat gpurun_out/r2d_tc3_roles.txt
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -25 gpurun_out/pytest_net3.log
timeout 600 python -m pytest tests/test_threads.py tests/test_gpu_encoders.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
