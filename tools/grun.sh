#!/bin/bash
# Retry wrapper around gpurun for "no slot right now" answers (exit 3 / transient): tools/grun.sh <timeout_s> '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out" | tail -12
  if echo "$out" | grep -q "status=transient"; then sleep 120; continue; fi
  exit $rc
done
exit 3
