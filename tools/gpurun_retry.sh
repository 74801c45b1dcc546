#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command string> [gpus]
# retries while the pod answers busy / transient (nothing is charged for those)
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|status=busy\|rc=3\|no box"; then echo "[retry $i] $(echo "$OUT" | grep gpurun | tail -2)"; sleep 120; continue; fi
  echo "$OUT"; exit 0
done
echo "gave up"; exit 3
