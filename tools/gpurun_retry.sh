#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
tag=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > gpurun_out/call_$tag.out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc after $i tries"; exit $rc; fi
  sleep 120
done
echo "gave up"; exit 3
