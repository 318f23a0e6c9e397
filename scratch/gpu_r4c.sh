#!/bin/bash
# leaner TMA epilogue (forward / data gradient): net + kernel parity on the tc3 engine, shape breakdowns, role accounting, bench
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r4c_pytest.log 2>&1; tail -n 5 gpurun_out/r4c_pytest.log
for w in pong navlaser navimg; do DDRL_PROF_SHAPES=1 timeout 300 python scratch/shape_prof.py $w > gpurun_out/r4c_shape_$w.txt 2>&1; head -n 12 gpurun_out/r4c_shape_$w.txt; done
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing_ps.so timeout 300 python scratch/tc3_roles_ps.py > gpurun_out/r4c_roles_ps.txt 2>&1; cat gpurun_out/r4c_roles_ps.txt
timeout 900 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r4c_bench.json 2> gpurun_out/r4c_bench.err; head -c 300 gpurun_out/r4c_bench.json; echo; tail -n 3 gpurun_out/r4c_bench.err
