#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'   -- retries while the pod answers busy (nothing charged)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 40 busy answers"
