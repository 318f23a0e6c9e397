#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest_gpu.log; tail -n 6 gpurun_out/r3f_pytest_gpu.log
timeout 900 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; head -c 300 gpurun_out/r3f_bench.json; echo; tail -n 3 gpurun_out/r3f_bench.err
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r3f_shape_pong.txt 2>&1; sed -n 12,30p gpurun_out/r3f_shape_pong.txt
