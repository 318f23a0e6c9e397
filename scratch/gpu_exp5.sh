#!/bin/bash
# A/B of a variant library: parity tests on the variant, then the per-shape profiles of both
LIB=$1
DDRL_LIB_PATH=ddrl4nav_b200/$LIB.so timeout -k 10 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in pong navimg navlaser; do
  echo "=== $wl product"; timeout -k 10 150 python scratch/shape_prof.py $wl 2>&1 | head -${2:-12}
  echo "=== $wl $LIB"; DDRL_LIB_PATH=ddrl4nav_b200/$LIB.so timeout -k 10 150 python scratch/shape_prof.py $wl 2>&1 | head -${2:-12}
done
