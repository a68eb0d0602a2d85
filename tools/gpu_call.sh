#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 600 python bench.py --no-cpu-baseline
run timeout 600 python bench.py --no-cpu-baseline
run timeout 300 python tools/prof_encoder.py
tail -5 $L
