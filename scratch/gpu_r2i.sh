#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_pytest_gpu.log 2>&1; tail -n 6 gpurun_out/r2i_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2i_bench_pong.json 2> gpurun_out/r2i_bench_pong.err; tail -c 4500 gpurun_out/r2i_bench_pong.json; tail -n 5 gpurun_out/r2i_bench_pong.err
