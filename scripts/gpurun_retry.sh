#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "transient/busy"
T=$1; shift
for i in $(seq 1 30); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT" | tail -40
  exit 0
done
echo "gave up after 30 transient answers"
exit 3
