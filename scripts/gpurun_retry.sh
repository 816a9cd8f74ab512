#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...] -- retries while the pod answers "busy" (rc 3)
LOG=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
