#!/bin/bash
# tools/gpu.sh <timeout_s> [--gpus N] <command> : gpurun with retries while the pod answers "transient" (exit 3)
T=$1; shift
EXTRA=""
if [ "$1" = "--gpus" ]; then EXTRA="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" $EXTRA -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
