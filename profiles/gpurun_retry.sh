#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> '<command>' [gpus]   — retries while the pod answers "transient/busy" (nothing is charged)
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  echo "$OUT" | tail -60
  if echo "$OUT" | grep -q "status=transient\|rc=3\|no box\|busy"; then echo "[retry $i] waiting 150 s"; sleep 150; continue; fi
  break
done
