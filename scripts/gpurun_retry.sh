#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy / transient)
tag=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$1" > "gpurun_out/${tag}_call.log" 2>&1
  rc=$?
  if grep -q "status=transient\|answers busy\|no box\|status=busy" "gpurun_out/${tag}_call.log" || [ $rc -eq 3 ]; then
    echo "attempt $i: busy ($(date +%H:%M:%S))" >> "gpurun_out/${tag}_retry.log"; sleep 100; continue
  fi
  break
done
echo "finished rc=$rc attempts=$i $(date +%H:%M:%S)" >> "gpurun_out/${tag}_retry.log"
