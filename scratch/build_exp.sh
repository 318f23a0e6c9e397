#!/bin/bash
# (historical: the T2_EXP / T2_NOFOLD / T2_ONE_STREAM / T2_SINGLE128 switches were removed from tc2.cu at the end of round 2 -- check out the round-1 tree to rebuild these experiments; results: profiles/r1d_tc2_skeleton.md)
# builds ddrl4nav_b200/libddrl_exp<mask>.so for each T2_EXP mask given (timing experiments; see csrc/tc2.cu)
cd "$(dirname "$0")/../ddrl4nav_b200/csrc" || exit 1
mkdir -p build_exp
OTHERS=$(ls build/*.o | grep -v '/tc2.o')
for e in "$@"; do
  ( nvcc -DT2_EXP=$e -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -c tc2.cu -o build_exp/tc2_$e.o &&
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libddrl_exp$e.so $OTHERS build_exp/tc2_$e.o -lcudart -lcuda ) &
done
wait
ls -la ../libddrl_exp*.so
