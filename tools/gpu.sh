#!/bin/bash
# retry wrapper around gpurun: exit code 3 = no box / slot free right now (nothing charged) -> wait and retry
# usage: tools/gpu.sh <timeout_s> [--gpus N] -- '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
