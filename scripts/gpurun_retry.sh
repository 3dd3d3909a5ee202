#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [gpus]  — retries while the pod answers "busy" (nothing charged)
t=$1; cmd=$2; gpus=${3:-1}
for attempt in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$cmd" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $t -- "$cmd" 2>&1); fi
  rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 120; continue; fi
  echo "$out" | tail -60
  exit $rc
done
echo "gave up after 40 busy answers"
