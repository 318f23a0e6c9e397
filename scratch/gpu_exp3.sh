#!/bin/bash
for g in 148 74 37; do for lib in libddrl_b200 libddrl_exp64; do echo "=== grid $g $lib"; DDRL_TC2_GRID=$g DDRL_LIB_PATH=ddrl4nav_b200/$lib.so timeout -k 10 120 python scratch/shape_prof.py pong 2>&1 | grep -E "conv_tc2\[fwd|gemm_tc2\[fwd|gemm_tc2\[dgrad"; done; done
