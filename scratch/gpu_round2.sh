#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_pong.json 2> gpurun_out/bench_pong.err; tail -c 4000 gpurun_out/bench_pong.json; tail -5 gpurun_out/bench_pong.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pong.csv python bench.py --profile-step > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:tc2 -c 20 -o gpurun_out/r1c_tc2 -f python bench.py --profile-step > gpurun_out/ncu_tc2.log 2>&1
ncu -i gpurun_out/r1c_tc2.ncu-rep --page raw --csv > gpurun_out/r1c_tc2_raw.csv 2>/dev/null
for f in gpurun_out/*.ncu-rep; do [ $(stat -c %s $f) -gt 30000000 ] && rm -f $f; done
du -sh gpurun_out
