#!/bin/sh
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>' : retry gpurun while the pod answers "transient" (nothing charged)
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -60
  echo "$out" | grep -q "status=transient" || exit 0
  sleep 45
done
exit 3
