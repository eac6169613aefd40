#!/bin/bash
# usage: tools/runs/submit.sh <timeout_s> <script> [log] [extra gpurun flags, e.g. --gpus 2]  -- gpurun with retries while the pod has no free slot (rc 3)
t=$1; s=$2; log=${3:-/tmp/gpurun_last.log}; shift; shift; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout $t -- "bash $s" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
