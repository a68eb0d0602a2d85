#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 600 python bench.py
run timeout 600 python bench.py --impl reference --steps 2 --warmup 1
run timeout 600 python bench.py --samples 128 --no-cpu-baseline --steps 10
run timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
tail -5 $L
