#!/bin/bash
# usage: tools/gpurun_retry_n.sh <gpus> <timeout> '<command>'  -- multi-GPU variant of gpurun_retry.sh
G=$1; T=$2; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
