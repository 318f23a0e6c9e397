#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest_gpu.log 2>&1; tail -n 25 gpurun_out/r2n_pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; head -c 700 gpurun_out/r2n_bench.json; tail -n 3 gpurun_out/r2n_bench.err
DDRL_PREP_LAUNCHES=1 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2n_bench_preplaunches.json 2> gpurun_out/r2n_bench_preplaunches.err; head -c 600 gpurun_out/r2n_bench_preplaunches.json
timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2n_shape_pong.txt 2>&1; head -40 gpurun_out/r2n_shape_pong.txt
