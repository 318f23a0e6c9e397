#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scratch/tc3_check.py > gpurun_out/tc3_check.txt 2>&1; cat gpurun_out/tc3_check.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; tail -15 gpurun_out/pytest_gemm.log
timeout 600 python -m pytest tests/test_threads.py tests/test_gpu_encoders.py -x -q > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
