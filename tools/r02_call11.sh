#!/bin/bash
# session-3 state check: full GPU suite (incl. local-radius tests), short bench
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call11}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu
run timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-step --no-reference-gpu
grep -n "^===\|^--- exit\|passed\|failed\|SUMMARY\|Error\|error" $L | cut -c1-220 | head -60
