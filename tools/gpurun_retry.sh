#!/bin/bash
# local helper (runs in the build container): submit a gpurun call, retrying while the pod answers "transient" / busy
# usage: tools/gpurun_retry.sh <timeout seconds> <log file> <command ...>
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|no box\|busy" $log && ! grep -q "exit code" $log; then sleep 90; continue; fi
  break
done
echo "gpurun_retry finished rc=$rc after $i attempt(s)" >> $log
