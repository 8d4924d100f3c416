#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> <command...> -- retries while the pod answers busy (exit 3 / transient)
T=$1; G=$2; shift 2
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/.retry.log 2>&1; fi
  rc=$?
  if grep -q "status=transient\|nothing was charged" gpurun_out/.retry.log; then sleep 90; continue; fi
  cat gpurun_out/.retry.log; exit $rc
done
cat gpurun_out/.retry.log; exit 3
