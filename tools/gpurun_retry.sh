#!/bin/bash
# usage: gpurun_retry.sh <gpus> <timeout> <command...> ; retries while the pod answers busy/transient
G=$1; T=$2; shift 2
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then
    echo "[retry $i] busy"; sleep 150; continue
  fi
  echo "$out" | tail -220
  exit 0
done
echo "gave up"
