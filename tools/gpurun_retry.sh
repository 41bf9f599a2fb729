#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>   — retries while the pod has no free slot (exit 3)
G=$1; T=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
cat /tmp/gpurun_last.log
