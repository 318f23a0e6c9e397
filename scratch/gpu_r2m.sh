#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -k "graph_replay or add_critic or learn_matches or micro_batching" > gpurun_out/r2m_pytest_new.log 2>&1; tail -n 30 gpurun_out/r2m_pytest_new.log
timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2m_bench_graph.json 2> gpurun_out/r2m_bench_graph.err; tail -c 1500 gpurun_out/r2m_bench_graph.json | cut -c1-1500; tail -n 3 gpurun_out/r2m_bench_graph.err
DDRL_NO_GRAPH=1 timeout 600 python bench.py --no-cpu --no-others --steps 4 --warmup 3 > gpurun_out/r2m_bench_nograph.json 2> gpurun_out/r2m_bench_nograph.err; head -c 600 gpurun_out/r2m_bench_nograph.json
