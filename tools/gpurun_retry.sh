#!/bin/bash
# gpurun_retry.sh <timeout> <script>: retry while the pod's GPU slots are busy (exit code 3), up to ~40 min
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
