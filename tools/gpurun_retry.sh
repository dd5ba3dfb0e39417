#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'  [extra gpurun flags...]
# Retries while the pod answers "busy" (exit code 3), at most 12 times, 2 minutes apart.
t=$1; cmd=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
