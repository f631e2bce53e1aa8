#!/bin/bash
# Retry wrapper around gpurun: exit code 3 (no slot right now, nothing charged) is retried every 45 s, up to ~40 min.
# usage: tools/grun.sh <timeout-seconds> '<command>' [extra gpurun flags]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 50); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
