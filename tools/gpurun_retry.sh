#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <command...>  -- gpurun, retried while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 120
done
tail -80 /tmp/gpurun_last.log
exit $rc
