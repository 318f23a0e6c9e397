#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or conv" > gpurun_out/r2v_pytest_k.log 2>&1; tail -n 3 gpurun_out/r2v_pytest_k.log
timeout 1200 python -m pytest tests/test_gpu_net.py tests/test_gpu_encoders.py -x -q > gpurun_out/r2v_pytest_net.log 2>&1; tail -n 3 gpurun_out/r2v_pytest_net.log
timeout 600 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; head -c 330 gpurun_out/r2v_bench.json; echo; tail -n 3 gpurun_out/r2v_bench.err
DDRL_TC3_FUSE_COLSUM=0 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2v_bench_nofuse.json 2> gpurun_out/r2v_bench_nofuse.err; head -c 330 gpurun_out/r2v_bench_nofuse.json; echo
