#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3b_pytest_gpu.log; tail -n 6 gpurun_out/r3b_pytest_gpu.log
timeout 900 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; head -c 300 gpurun_out/r3b_bench.json; echo; tail -n 3 gpurun_out/r3b_bench.err
DDRL_DETERMINISTIC=1 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r3b_bench_det.json 2> gpurun_out/r3b_bench_det.err; head -c 300 gpurun_out/r3b_bench_det.json; echo
