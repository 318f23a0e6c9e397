#!/bin/bash
# sign-bit activation masks (forward launches write them, data gradients read them)
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4m_pytest.log 2>&1; tail -n 5 gpurun_out/r4m_pytest.log
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4m_shape_pong.txt 2>&1; head -n 14 gpurun_out/r4m_shape_pong.txt
DDRL_NO_SIGNBITS=1 DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4m_shape_pong_nobits.txt 2>&1; head -n 14 gpurun_out/r4m_shape_pong_nobits.txt
timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r4m_bench.json 2> gpurun_out/r4m_bench.err; head -c 300 gpurun_out/r4m_bench.json; echo; tail -n 3 gpurun_out/r4m_bench.err
