#!/bin/bash
# SASS opcode histogram of the engine objects (cuobjdump -sass): the tcgen05 / TMA / TMEM mnemonics that prove the path
#   UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
#   SYNCS = mbarrier ops, UTCCP = tcgen05.cp
cd "$(dirname "$0")/../ddrl4nav_b200/csrc/build" || exit 1
for f in tc3 tc2 gemm_tc prep easybytes gae ppo_loss adam heads; do
  echo "## $f.o  (kernels: $(cuobjdump -sass $f.o | grep -c 'Function :'))"
  cuobjdump -sass $f.o | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/;$//' \
    | awk -F. '{k=$1; if (k ~ /^(UTCHMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCCP|SYNCS|LDG|STG|LDS|STS|RED|ATOMG|ATOM)$/ && NF>1) k=$1"."$2; c[k]++} END{for (k in c) print c[k], k}' \
    | sort -rn > /tmp/sass_hist.$$
  echo -n "blackwell ops:  "; grep -E ' (UTC|UTMA|LDTM|STTM|SYNCS|ELECT|FENCE|UBLKCP|UBLKRED|CCTL)' /tmp/sass_hist.$$ | awk '{printf "%s:%s  ", $2, $1} END{print ""}'
  echo -n "top opcodes:    "; head -40 /tmp/sass_hist.$$ | awk '{printf "%s:%s  ", $2, $1} END{print ""}'; rm -f /tmp/sass_hist.$$
done
