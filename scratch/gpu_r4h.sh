#!/bin/bash
# halo boxes in the pre-split first-conv weight gradient
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 600 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "presplit or variants or (pong and (golden or oracle or learn or determin))" > gpurun_out/r4h_pytest.log 2>&1; tail -n 12 gpurun_out/r4h_pytest.log
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4h_shape_pong.txt 2>&1; head -n 8 gpurun_out/r4h_shape_pong.txt
timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err; head -c 300 gpurun_out/r4h_bench.json; echo; tail -n 3 gpurun_out/r4h_bench.err
