#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (exit code 3: nothing charged)
# usage: scripts/gpurun_retry.sh <log> <timeout> <command...>
log=$1; shift; to=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
