#!/bin/bash
# gpu.sh <timeout_s> <log> <command...>: run a command on the B200 box through gpurun, retrying while the pod is busy.
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -eq 3 ] || grep -q "status=transient" "$LOG"; then sleep 45; continue; fi
  break
done
echo "[gpu.sh] done rc=$rc" >> "$LOG"
