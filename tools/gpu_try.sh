#!/bin/bash
# usage: tools/gpu_try.sh <timeout-seconds> '<command>' [gpus]   — retries while the pod answers busy/transient
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$out" | grep -q "status=transient\|nothing was charged\|status=busy"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up"
