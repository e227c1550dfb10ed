#!/bin/bash
# usage: gpu_retry.sh [-g N] <timeout_s> <command string>   -- retries while the pod answers "busy" (exit 3)
g=""
if [ "$1" = "-g" ]; then g="--gpus $2"; shift 2; fi
t=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $g --timeout $t -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
