#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> '<command>'   -- retries while the pod is busy (rc 3)
TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 60; continue; fi
  break
done
cat /tmp/gpurun_last.log | tail -80
