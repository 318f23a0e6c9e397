#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout> [--gpus N] -- '<command>'   : retries while the pod answers busy (exit 3 / transient)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun: gave up after 40 busy answers"; exit 3
