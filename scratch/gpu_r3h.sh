#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py tests/test_easybytes.py tests/test_threads.py -x -q -m gpu > gpurun_out/r3h_pytest_net.log 2>&1; tail -n 4 gpurun_out/r3h_pytest_net.log
timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r3h_bench.json 2> gpurun_out/r3h_bench.err; head -c 300 gpurun_out/r3h_bench.json; echo; tail -n 3 gpurun_out/r3h_bench.err
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r3h_shape_pong.txt 2>&1; grep -E "s2d|amax|one learn" gpurun_out/r3h_shape_pong.txt
