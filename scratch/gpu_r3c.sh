#!/bin/bash
# line-level ncu profile of the fused parity data gradient of Pong's conv2 (tc3_kernel<128>, the 10th tc3_kernel launch of an iteration)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:^tc3_kernel --launch-skip 9 --launch-count 1 -o gpurun_out/r3c_dgrad -f python bench.py --profile-step > gpurun_out/r3c_ncu.log 2>&1; tail -n 3 gpurun_out/r3c_ncu.log
python scratch/ncu_lines.py gpurun_out/r3c_dgrad.ncu-rep 60 > gpurun_out/r3c_dgrad_lines.txt 2>&1; head -n 70 gpurun_out/r3c_dgrad_lines.txt | cut -c1-260
ncu -i gpurun_out/r3c_dgrad.ncu-rep --page raw --csv > gpurun_out/r3c_dgrad_raw.csv 2>/dev/null
ls -la gpurun_out/r3c_dgrad.ncu-rep
