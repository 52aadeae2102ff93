#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command...>   -- retries while the pod answers "transient" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  break
done
