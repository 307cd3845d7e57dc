#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free, nothing charged).  Usage: scripts/gpurun_retry.sh <timeout_s> '<command>' [gpus]
t=$1; cmd=$2; gpus=${3:-1}
for i in $(seq 1 12); do
  if [ "$gpus" = "1" ]; then out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$cmd" 2>&1); else out=$(/usr/local/graft/bin/gpurun --gpus $gpus --timeout $t -- "$cmd" 2>&1); fi
  rc=$?
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] transient"; sleep 45; continue; fi
  echo "$out" | tail -40
  exit $rc
done
echo "gave up"; exit 3
