#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers busy (exit 3, nothing charged)
LOG=$1; shift; TMO=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 90
done
echo "finished rc=$rc tries=$i" >> $LOG
