#!/bin/bash
# retries a gpurun call while the pod answers "busy / draining" (exit code 3 or status=transient, nothing charged)
#   tools/gpu_submit.sh [gpurun options] -- 'command'
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then
    echo "[gpu_submit] attempt $attempt: busy, retrying in 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit $rc
done
echo "[gpu_submit] gave up after 30 attempts"; exit 3
