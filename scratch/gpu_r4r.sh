#!/bin/bash
# non-reference 3 x 960 laser variant: parity against the oracle + bench line; regression of the other nets' gradient tests
set -x
mkdir -p gpurun_out
DDRL_TEST_GEMM_MODE=tc3 timeout 200 python -m pytest tests/test_gpu_net.py -x -q -m gpu -k "grads_match_oracle or (forward_matches and nav)" > gpurun_out/r4r_pytest.log 2>&1; tail -n 3 gpurun_out/r4r_pytest.log
timeout 150 python bench.py --workload navlaser3 --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r4r_bench_navlaser3.json 2> gpurun_out/r4r_bench.err; head -c 600 gpurun_out/r4r_bench_navlaser3.json; echo; tail -n 2 gpurun_out/r4r_bench.err
