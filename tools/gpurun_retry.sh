#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (status=transient, nothing charged).
#   tools/gpurun_retry.sh [--gpus N] <timeout_s> '<command>'
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$OUT"
  exit 0
done
echo "gpurun_retry: gave up after 40 attempts"
exit 3
