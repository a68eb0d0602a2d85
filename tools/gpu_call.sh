#!/bin/bash
# One gpurun session (edit per experiment): validation + A/B timings + ncu capture.  Run from the repo root on the GPU box.
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 300 python tools/prof_encoder.py
run timeout 600 python bench.py
tail -5 $L
