#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   -- retries while gpurun answers "no box / slot free" (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 75
done
exit 3
