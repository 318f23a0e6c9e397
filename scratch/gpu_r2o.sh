#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_net.py -x -q > gpurun_out/r2o_pytest_net.log 2>&1; tail -n 4 gpurun_out/r2o_pytest_net.log
timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; head -c 400 gpurun_out/r2o_bench.json; tail -n 3 gpurun_out/r2o_bench.err
for k in pong navlaser navimg; do timeout 300 python scratch/shape_prof.py $k > gpurun_out/r2o_shape_$k.txt 2>&1; head -12 gpurun_out/r2o_shape_$k.txt; done
