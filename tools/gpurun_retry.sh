#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers "transient / busy" (nothing is charged then)
for i in $(seq 1 24); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 45
done
