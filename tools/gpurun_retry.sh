#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] -- <command>   (retries while the pod answers "busy": exit code 3)
log=$1; shift; to=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
