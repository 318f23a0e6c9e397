#!/bin/bash
# conversion-task offsets precomputed (all weight gradients); halo boxes on / off
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4i_pytest.log 2>&1; tail -n 3 gpurun_out/r4i_pytest.log
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4i_shape_pong.txt 2>&1; head -n 12 gpurun_out/r4i_shape_pong.txt
DDRL_TC3_NO_HALO=1 DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4i_shape_pong_nohalo.txt 2>&1; head -n 6 gpurun_out/r4i_shape_pong_nohalo.txt
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py navlaser > gpurun_out/r4i_shape_navlaser.txt 2>&1; head -n 8 gpurun_out/r4i_shape_navlaser.txt
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py navimg > gpurun_out/r4i_shape_navimg.txt 2>&1; head -n 6 gpurun_out/r4i_shape_navimg.txt
