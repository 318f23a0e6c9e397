#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv or gemm" > gpurun_out/r3e_pytest_k.log 2>&1; tail -n 4 gpurun_out/r3e_pytest_k.log
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py -x -q > gpurun_out/r3e_pytest_net.log 2>&1; tail -n 6 gpurun_out/r3e_pytest_net.log
timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; head -c 300 gpurun_out/r3e_bench.json; echo; tail -n 3 gpurun_out/r3e_bench.err
DDRL_TC3_TMA_DGRAD=0 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r3e_bench_off.json 2>/dev/null; head -c 300 gpurun_out/r3e_bench_off.json; echo
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r3e_shape_pong.txt 2>&1; head -n 16 gpurun_out/r3e_shape_pong.txt
