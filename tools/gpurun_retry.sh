#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [gpurun args...] -- retries while the pod answers "busy" (exit code 3, nothing charged)
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
