#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout-seconds> [--gpus N] -- <command...>
# Retries while the pod answers "busy" (exit code 3 / transient), up to ~40 min.
LOG=$1; TO=$2; shift 2
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun rc=$rc attempts=$attempt" >> $LOG
