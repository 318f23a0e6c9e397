#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/pytest_k3.log 2>&1; tail -5 gpurun_out/pytest_k3.log
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -8 gpurun_out/pytest_net3.log
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2f_shape_pong_tc3.txt 2>&1; head -20 gpurun_out/r2f_shape_pong_tc3.txt
DDRL_TC3_NO_TMA_STORE=1 DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2f_shape_pong_tc3_nostore.txt 2>&1; head -12 gpurun_out/r2f_shape_pong_tc3_nostore.txt
