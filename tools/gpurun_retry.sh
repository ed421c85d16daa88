#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  Usage:
#   tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun finished rc=$rc after $i tries" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up" >> "$log"; exit 3
