#!/bin/bash
# gp.sh <tag> <timeout-s> <command...>: gpurun with retries while the pod answers "transient/busy" (nothing charged)
tag=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" gpurun_out/${tag}_call.log || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
tail -2 gpurun_out/${tag}_call.log
