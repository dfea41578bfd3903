#!/bin/bash
# usage: profiles/gpurun_retry.sh <timeout_s> '<command>' — retries while the pod answers "transient" (nothing charged).
t=$1; shift
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gave up after 40 transient answers"
