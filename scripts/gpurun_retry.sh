#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <timeout> <command...>   -- retries while the pod answers "transient"
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun ${GPURUN_ARGS:-} --timeout $to -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 90; else break; fi
done
