#!/bin/bash
# gpurun_retry.sh <log> <gpurun args...>: retries while the pod answers "busy" (exit code 3; nothing is charged)
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
