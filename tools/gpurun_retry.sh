#!/bin/bash
# developer helper: retry a gpurun call while the pod answers "no slot free" (exit code 3), at most 20 times
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 150
done
exit 3
