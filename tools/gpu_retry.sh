#!/bin/bash
# developer tool: run a gpurun call, retrying while the pod answers "busy" (exit 3).  usage: tools/gpu_retry.sh <log> <timeout> <command...>
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; rc=$?
  [ $rc -ne 3 ] && break
  sleep 120
done
echo "gpurun rc=$rc" >> $log
