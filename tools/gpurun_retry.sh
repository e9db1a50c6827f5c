#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   - retries while the pod answers busy / transient (nothing charged)
t=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|exit code 3\|status=busy"; then sleep 60; continue; fi
  break
done
