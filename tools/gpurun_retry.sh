#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3; nothing is charged for those).  Usage:
#   [GPUS=2] tools/gpurun_retry.sh <log> <timeout-seconds> '<command>'
log=$1; to=$2; shift 2
extra=""
if [ -n "${GPUS:-}" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $extra --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$log"; exit 3
