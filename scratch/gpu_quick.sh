#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
for wl in pong navimg navlaser; do timeout -k 10 200 python scratch/shape_prof.py $wl > gpurun_out/q_${wl}.txt 2>&1; head -${1:-16} gpurun_out/q_${wl}.txt; done
