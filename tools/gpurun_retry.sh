#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   - retries while the pod answers busy / transient (nothing charged)
# GPURUN_ATTEMPTS (default 12) bounds the retries; each refused attempt costs up to ~4 minutes of waiting
t=$1; shift
for i in $(seq 1 ${GPURUN_ATTEMPTS:-12}); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|exit code 3\|status=busy"; then sleep 45; continue; fi
  break
done
