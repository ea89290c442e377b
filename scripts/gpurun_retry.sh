#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit code 3)
T=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 60
done
exit 3
