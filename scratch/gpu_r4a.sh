#!/bin/bash
# pre-split first conv (SS MMAs): parity of the Pong net tests, shape breakdown, bench line
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 600 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "presplit or variants or (pong and (golden or oracle or learn))" > gpurun_out/r4a_pytest.log 2>&1; tail -n 15 gpurun_out/r4a_pytest.log
DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4a_shape_pong.txt 2>&1; head -n 14 gpurun_out/r4a_shape_pong.txt; grep -E "presplit|s2d" gpurun_out/r4a_shape_pong.txt
DDRL_NO_PRESPLIT_WGRAD=1 DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py pong > gpurun_out/r4a_shape_pong_nowg.txt 2>&1; head -n 6 gpurun_out/r4a_shape_pong_nowg.txt
timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; head -c 300 gpurun_out/r4a_bench.json; echo; tail -n 5 gpurun_out/r4a_bench.err
