#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   — retries while the pod has no free slot (nothing is charged then)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|no box\|retry in a few minutes" "$LOG"; then sleep 90; continue; fi
  break
done
