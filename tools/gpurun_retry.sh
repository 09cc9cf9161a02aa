#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged).  usage: tools/gpurun_retry.sh [gpurun args...] -- 'cmd'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
