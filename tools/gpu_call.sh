#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu
tail -5 $L
