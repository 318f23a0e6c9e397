#!/bin/bash
# round 2, call A: tensor-core issue-rate microbenchmarks + cuBLAS reference rates + the current tree's shape profile
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 300 ./scratch/mma_bench > gpurun_out/mma_bench.json 2> gpurun_out/mma_bench.err; cat gpurun_out/mma_bench.json
timeout 300 python scratch/cublas_peaks.py > gpurun_out/cublas_peaks.json 2> gpurun_out/cublas_peaks.err; cat gpurun_out/cublas_peaks.json
timeout 300 python scratch/shape_prof.py pong > gpurun_out/r2a_shape_pong.txt 2>&1; tail -40 gpurun_out/r2a_shape_pong.txt
