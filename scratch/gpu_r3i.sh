#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py -x -q -m gpu > gpurun_out/r3i_pytest_net.log 2>&1; tail -n 4 gpurun_out/r3i_pytest_net.log
timeout 900 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err; head -c 300 gpurun_out/r3i_bench.json; echo; tail -n 3 gpurun_out/r3i_bench.err
for w in navlaser navimg; do DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py $w > gpurun_out/r3i_shape_$w.txt 2>&1; head -n 22 gpurun_out/r3i_shape_$w.txt; done
