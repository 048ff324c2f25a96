#!/bin/bash
# usage: scripts/grun.sh <gpus> <timeout_s> <command...> : retries while the pod has no free slot (nothing is charged)
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    echo "[grun] attempt $i: pod busy, retrying in 90 s" >&2
    sleep 90
    continue
  fi
  echo "$out"
  exit 0
done
echo "[grun] gave up"; exit 3
