#!/bin/bash
# gpu_retry.sh <timeout> <log> <command...>: run gpurun, retrying while the pod answers "busy" (exit code 3).
# GPURUN_FLAGS (environment) is placed before `--`, e.g. GPURUN_FLAGS="--gpus 2".
T=$1; LOG=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
