#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/r2u_pytest_k.log 2>&1; tail -n 3 gpurun_out/r2u_pytest_k.log
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py -x -q > gpurun_out/r2u_pytest_net.log 2>&1; tail -n 3 gpurun_out/r2u_pytest_net.log
timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; head -c 330 gpurun_out/r2u_bench.json; echo; tail -n 3 gpurun_out/r2u_bench.err
DDRL_TC3_TMA_STORE=0 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2u_bench_vecstore.json 2> gpurun_out/r2u_bench_vecstore.err; head -c 330 gpurun_out/r2u_bench_vecstore.json; echo
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing.so timeout 600 python scratch/tc3_roles.py > gpurun_out/r2u_tc3_roles.txt 2>&1; head -n 36 gpurun_out/r2u_tc3_roles.txt
for k in pong navlaser navimg; do timeout 300 python scratch/shape_prof.py $k > gpurun_out/r2u_shape_$k.txt 2>&1; head -8 gpurun_out/r2u_shape_$k.txt; done
