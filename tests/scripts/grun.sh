#!/bin/bash
# retry wrapper around gpurun for "no slot right now" (exit code 3): usage grun.sh <timeout-s> <logfile> <command string>
t=$1; log=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
