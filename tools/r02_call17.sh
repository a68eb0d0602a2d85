#!/bin/bash
# final validation of the session-3 state: full GPU suite, smoke, bench (quick)
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call17}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 600 python bench.py --steps 20 --warmup 3 --quick
grep -n "^===\|^--- exit\|passed\|failed\|smoke:\|Error" $L | cut -c1-220
