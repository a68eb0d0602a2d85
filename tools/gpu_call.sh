#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unfused_api.py tests/test_gpu_model.py -q -m gpu -k "gather or render or unfused or forward or full_size"
run timeout 300 python tools/prof_kernels.py --which gather --reps 10
MNF_GATHER_OCC=6 run timeout 300 python tools/prof_kernels.py --which gather --reps 10
run timeout 300 python tools/prof_kernels.py --which gather --reps 10 --rays 327680
tail -5 $L
