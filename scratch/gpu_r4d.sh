#!/bin/bash
# store warp + pipelined panel loop + bias prefetch
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4f_pytest.log 2>&1; tail -n 3 gpurun_out/r4f_pytest.log
for w in pong navlaser; do DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py $w > gpurun_out/r4f_shape_$w.txt 2>&1; head -n 8 gpurun_out/r4f_shape_$w.txt; done
grep "K=64,\|K=96," gpurun_out/r4f_shape_navlaser.txt
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing_ps.so timeout 300 python scratch/tc3_roles_ps.py > gpurun_out/r4f_roles_ps.txt 2>&1; cat gpurun_out/r4f_roles_ps.txt
timeout 900 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err; head -c 300 gpurun_out/r4f_bench.json; echo; tail -n 3 gpurun_out/r4f_bench.err
