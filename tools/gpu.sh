#!/bin/bash
# tools/gpu.sh LOG TIMEOUT [--gpus N] -- 'command' : gpurun with retries while the pod's GPU slots are busy (exit 3)
log=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
