#!/bin/bash
# gpurun_retry.sh <timeout> <script> [extra gpurun options, e.g. --gpus 2]: retry while the pod's GPU slots are busy
t=$1; s=$2; shift 2
for i in $(seq 1 25); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "bash $s"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
