#!/bin/bash
# gpurun_retry.sh <log> <timeout> <command...>: retry while the pod answers busy (exit 3 / "transient")
log=$1; shift; to=$1; shift
for i in $(seq 1 30); do
  gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient\|rc=3" $log; then exit 0; fi
  sleep 150
done
exit 3
