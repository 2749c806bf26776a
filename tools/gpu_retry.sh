#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- <command>   : gpurun with retries while the pod answers busy; log in gpurun_out/<tag>_gpurun.log
tag=$1; to=$2; shift 2
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift 2; fi
shift   # the --
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > gpurun_out/${tag}_gpurun.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${tag}_gpurun.log; then break; fi
  sleep 90
done
echo "done rc=$rc" >> gpurun_out/${tag}_gpurun.log
