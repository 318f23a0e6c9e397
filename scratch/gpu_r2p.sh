#!/bin/bash
set -x
mkdir -p gpurun_out
DDRL_LIB_PATH=ddrl4nav_b200/libddrl_b200_timing.so timeout 600 python scratch/tc3_roles.py > gpurun_out/r2p_tc3_roles.txt 2>&1; tail -n 30 gpurun_out/r2p_tc3_roles.txt
