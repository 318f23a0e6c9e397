#!/bin/bash
# two-phase K tiling of the weight gradient (small maps)
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4l_pytest.log 2>&1; tail -n 3 gpurun_out/r4l_pytest.log
for w in pong navlaser navimg; do DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py $w > gpurun_out/r4l_shape_$w.txt 2>&1; head -n 1 gpurun_out/r4l_shape_$w.txt; grep wgrad gpurun_out/r4l_shape_$w.txt | head -8; done
timeout 900 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r4l_bench.json 2> gpurun_out/r4l_bench.err; head -c 300 gpurun_out/r4l_bench.json; echo; tail -n 3 gpurun_out/r4l_bench.err
