#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>  -- retries while the pod answers "busy" (exit 3)
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
