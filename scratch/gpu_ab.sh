#!/bin/bash
# A/B of engine knobs on the per-shape step profile (one learn step each)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for wl in pong navimg navlaser; do
  timeout 200 python scratch/shape_prof.py $wl > gpurun_out/ab_${wl}_default.txt 2>&1
  DDRL_TC2_CONV_MAXBN=64 timeout 200 python scratch/shape_prof.py $wl > gpurun_out/ab_${wl}_maxbn64.txt 2>&1
  DDRL_TC2_CONV_MAXBN=64 DDRL_TC2_CONV_MAXBN_KB=16 timeout 200 python scratch/shape_prof.py $wl > gpurun_out/ab_${wl}_maxbn64_kb16.txt 2>&1
done
DDRL_NO_S2D=1 timeout 200 python scratch/shape_prof.py pong > gpurun_out/ab_pong_nos2d.txt 2>&1
head -14 gpurun_out/ab_pong_default.txt
head -3 gpurun_out/ab_*_maxbn64*.txt gpurun_out/ab_navimg_default.txt gpurun_out/ab_navlaser_default.txt
