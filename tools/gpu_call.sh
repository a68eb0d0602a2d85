#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
export MNF_GATHER_IMPL=4
run timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gather or pack or render"
run timeout 600 python -m pytest tests/test_gpu_unfused_api.py tests/test_gpu_model.py -q -m gpu
run timeout 300 python tools/prof_kernels.py --which gather --reps 5
tail -5 $L
