#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <timeout_s> <gpus> <command...>   -- retries while the pod answers "busy" (exit code 3 / transient)
log=$1; to=$2; gpus=$3; shift 3
for i in $(seq 1 12); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $log 2>&1; fi
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
