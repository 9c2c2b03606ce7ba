#!/bin/bash
# tools/gpu.sh <timeout_s> <command...> : gpurun with retries while the pod answers "transient" (exit 3)
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
