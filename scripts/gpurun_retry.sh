#!/bin/bash
# usage: gpurun_retry.sh <gpurun args...>
# Retries ONLY while gpurun answers "busy" with nothing charged (exit 3 and no strike); a lost box is never retried blindly.
for i in $(seq 1 40); do
  out=$(mktemp)
  /usr/local/graft/bin/gpurun "$@" 2>&1 | tee "$out"
  rc=${PIPESTATUS[0]}
  if [ $rc -ne 3 ]; then rm -f "$out"; exit $rc; fi
  if grep -qi "strike\|lost\|stopped responding" "$out"; then rm -f "$out"; echo "[retry] box lost: NOT retrying"; exit 4; fi
  rm -f "$out"
  sleep 45
done
exit 3
