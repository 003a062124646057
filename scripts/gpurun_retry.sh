#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <command...>   -- retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
tail -40 /tmp/gpurun_last.log
