#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <script> [gpus]: retries while the pod answers busy (exit 3 / transient)
T=$1; S=$2; G=${3:-1}
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  if [ "$G" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S" > /tmp/gpurun_last.log 2>&1
  fi
  if grep -q "status=transient\|nothing was charged" /tmp/gpurun_last.log; then sleep 90; continue; fi
  break
done
tail -70 /tmp/gpurun_last.log | cut -c1-1800
