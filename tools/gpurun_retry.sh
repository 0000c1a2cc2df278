#!/bin/bash
# gpurun with retries while the pod's GPU slots are busy (exit code 3 = nothing charged)
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 60
done
exit 3
