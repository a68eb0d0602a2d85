#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "decoder or render or forward or full_size or blender"
for i in 1 2; do
run timeout 300 python tools/prof_kernels.py --which decoder --impl 2 --reps 10
MNF_LIB_PATH=$PWD/matchnerf_b200/variants/lib_base.so run timeout 300 python tools/prof_kernels.py --which decoder --impl 2 --reps 10
done
tail -5 $L
