#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:s2d_c4 --launch-count 1 -o gpurun_out/r3g_s2d -f python bench.py --profile-step > gpurun_out/r3g_ncu.log 2>&1; tail -n 2 gpurun_out/r3g_ncu.log
python scratch/ncu_lines.py gpurun_out/r3g_s2d.ncu-rep 25 2>&1 | cut -c1-230
ncu -i gpurun_out/r3g_s2d.ncu-rep --page details 2>/dev/null | grep -E "Duration|DRAM Throughput|Memory Throughput|L1/TEX Hit|L2 Hit|Achieved Occupancy|Theoretical Occupancy|Registers Per|Shared Memory Config|Bank conflicts|Issue Slots Busy|Executed Ipc|Max Bandwidth|Mem Busy|uncoalesced|excessive|sectors" | head -30
