#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>: retries while the pod answers "busy" (nothing is charged for those)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_ARGS --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then exit 0; fi
  sleep 120
done
