#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers "transient/busy"
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|status=busy\|no box\|rc=3"; then sleep 60; continue; fi
  echo "$OUT" | tail -40
  exit 0
done
echo "gave up after 40 transient answers"
exit 3
