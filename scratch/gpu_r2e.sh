#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/pytest_net3.log 2>&1; tail -8 gpurun_out/pytest_net3.log
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2e_shape_pong_tc3.txt 2>&1; head -24 gpurun_out/r2e_shape_pong_tc3.txt
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py navlaser > gpurun_out/r2e_shape_navlaser_tc3.txt 2>&1; head -30 gpurun_out/r2e_shape_navlaser_tc3.txt
DDRL_GEMM_MODE=tc3 timeout 300 python scratch/shape_prof.py navimg > gpurun_out/r2e_shape_navimg_tc3.txt 2>&1; head -30 gpurun_out/r2e_shape_navimg_tc3.txt
timeout 600 python -m pytest tests/test_threads.py tests/test_gpu_encoders.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -5 gpurun_out/pytest_new.log
