#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "=== product"; timeout 120 python scratch/tc2_exp.py 2>&1 | tail -16
for e in "$@"; do echo "=== exp $e"; DDRL_LIB_PATH=ddrl4nav_b200/libddrl_exp$e.so timeout 120 python scratch/tc2_exp.py 2>&1 | tail -16; done
for wl in navimg navlaser; do timeout 200 python scratch/shape_prof.py $wl > gpurun_out/ab2_${wl}.txt 2>&1; head -8 gpurun_out/ab2_${wl}.txt; done
