#!/bin/bash
# tools/gpu_retry.sh <timeout_s> <script> <log>: run a script on the GPU box, retrying while the pod is busy
T=$1; S=$2; L=$3; shift 3
for i in $(seq 1 30); do
  gpurun --timeout $T "$@" -- bash $S > $L 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" $L || grep -q "retry in a few minutes" $L; then sleep 90; else break; fi
done
