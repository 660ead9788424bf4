#!/bin/bash
# Local helper (not for the GPU box): retry a gpurun call while the pod answers "busy / draining" (exit 3).
#   bash tools/gpurun_retry.sh <out-file> [--gpus N] [--timeout S] -- '<command>'
OUT=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$OUT"; then exit $rc; fi
  sleep 120
done
exit 3
