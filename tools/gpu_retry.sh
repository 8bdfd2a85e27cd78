#!/bin/bash
# gpurun with retry while the pod has no free GPU slot (exit code 3 = nothing charged).  usage: tools/gpu_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
