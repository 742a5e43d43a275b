#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command...>   -- retries while the pod has no free slot (exit 3), up to ~1 h
T=$1; shift
for i in $(seq 1 30); do
    /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 60
done
exit 3
