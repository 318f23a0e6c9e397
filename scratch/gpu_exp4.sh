#!/bin/bash
for e in "$@"; do echo "=== exp $e"; DDRL_LIB_PATH=ddrl4nav_b200/libddrl_exp$e.so timeout -k 10 120 python scratch/shape_prof.py pong 2>&1 | grep -E "conv_tc2\[fwd|gemm_tc2\[fwd|gemm_tc2\[dgrad"; done
