#!/bin/bash
mkdir -p gpurun_out
for e in "$@"; do echo "=== exp $e"; DDRL_LIB_PATH=ddrl4nav_b200/libddrl_exp$e.so timeout -k 10 120 python scratch/shape_prof.py pong 2>&1 | head -13; done
