#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc3_wgrad_kernel -s 2 -c 1 -o gpurun_out/r2r_tc3_wgrad_conv2 -f python scratch/one_conv.py tc3 2 8192 20 20 32 64 4 4 2 0 > gpurun_out/r2r_ncu.log 2>&1; tail -3 gpurun_out/r2r_ncu.log
ncu -i gpurun_out/r2r_tc3_wgrad_conv2.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r2r_wgrad_source.csv 2>/dev/null; wc -l gpurun_out/r2r_wgrad_source.csv
ls -la gpurun_out/*.ncu-rep
