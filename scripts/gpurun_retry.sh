#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> [--gpus N] -- '<command>'   (retries while the pod answers busy / transient)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy\|no box\|retry in a few minutes" $LOG && ! grep -q "status=ok" $LOG; then
    sleep 60; continue
  fi
  break
done
echo "gpurun_retry done rc=$rc" >> $LOG
