#!/bin/bash
# usage: gpurun_retry.sh <gpurun args...>   -- retries while gpurun answers "busy" (exit 3), nothing is charged for those
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
