#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box / slot free; nothing is charged for those)
# usage: tools/gpurun_retry.sh <max tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] transient"; sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "gave up after $tries transient answers"; exit 3
