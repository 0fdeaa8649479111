#!/bin/bash
# usage: tools/grun.sh <logname> <timeout-seconds> '<command>'   -- gpurun with retries while the pod has no free slot (exit 3)
log=gpurun_out/$1.log; to=$2; shift 2
mkdir -p gpurun_out
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
exit 3
