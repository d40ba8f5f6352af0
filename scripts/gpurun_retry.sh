#!/bin/bash
# gpurun with retries while the pod answers busy/transient (nothing charged):  scripts/gpurun_retry.sh <timeout_s> '<command>'
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1); rc=$?
  echo "$out" | tail -120
  if echo "$out" | grep -q "status=transient\|retry in a few minutes\|no box\|busy"; then echo "[retry $i] sleeping 120 s"; sleep 120; continue; fi
  exit $rc
done
exit 3
