#!/bin/bash
# compute-sanitizer passes over the kernel tests and the tc3 net tests (memcheck: every kernel; racecheck / synccheck: the
# kernels that synchronise through shared memory, and the tcgen05 engine for what the tools understand of mbarrier / TMA)
set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
export DDRL_TEST_GEMM_MODE=tc3
timeout 1200 $S --tool memcheck --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_kernels.py -x -q -k "not full_size" > gpurun_out/r2s_memcheck_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_memcheck_kernels.log; tail -n 6 gpurun_out/r2s_memcheck_kernels.log
timeout 1200 $S --tool memcheck --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_net.py tests/test_easybytes.py -x -q -k "forward_matches or backward_grads or learn_matches or segments or easybytes or device" > gpurun_out/r2s_memcheck_net.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_memcheck_net.log; tail -n 6 gpurun_out/r2s_memcheck_net.log
timeout 900 $S --tool racecheck --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_kernels.py -x -q -k "(gae or loss or adam or head or sampl) and not full_size" > gpurun_out/r2s_racecheck_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_racecheck_kernels.log; tail -n 6 gpurun_out/r2s_racecheck_kernels.log
timeout 900 $S --tool synccheck --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_kernels.py -x -q -k "not full_size" > gpurun_out/r2s_synccheck_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_synccheck_kernels.log; tail -n 6 gpurun_out/r2s_synccheck_kernels.log
timeout 900 $S --tool racecheck --error-exitcode 77 --print-limit 30 python -m pytest tests/test_gpu_kernels.py -x -q -k "tc3 or gemm or conv" > gpurun_out/r2s_racecheck_tc3.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_racecheck_tc3.log; tail -n 12 gpurun_out/r2s_racecheck_tc3.log
