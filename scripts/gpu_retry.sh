#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
#   scripts/gpu_retry.sh <timeout_s> <log> [--gpus N] -- '<command>'
T=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
