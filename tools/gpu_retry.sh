#!/bin/bash
# usage: [GPUS=N] gpu_retry.sh <timeout-seconds> <command...> : re-submits while the pod answers "transient / busy" (exit 3)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?
  else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1; rc=$?; fi
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
cat /tmp/gpurun_last.log | tail -60
