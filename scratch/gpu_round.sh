#!/bin/bash
# One dense GPU call: parity tests, three bench lines, ncu launch list, ncu full capture of the tc2 engine and GAE.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_pong.json 2> gpurun_out/bench_pong.err; tail -c 3000 gpurun_out/bench_pong.json
timeout 600 python bench.py --workload navlaser > gpurun_out/bench_navlaser.json 2> gpurun_out/bench_navlaser.err
timeout 600 python bench.py --workload navimg > gpurun_out/bench_navimg.json 2> gpurun_out/bench_navimg.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_pong.csv python bench.py --profile-step > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc2 -c 24 -o gpurun_out/r1_tc2 -f python bench.py --profile-step > gpurun_out/ncu_tc2.log 2>&1
timeout 300 python scratch/shape_prof.py pong > gpurun_out/shape_pong.txt 2>&1
timeout 300 python scratch/shape_prof.py navlaser > gpurun_out/shape_navlaser.txt 2>&1
timeout 300 python scratch/shape_prof.py navimg > gpurun_out/shape_navimg.txt 2>&1
ls -la gpurun_out
