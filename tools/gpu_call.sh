#!/bin/bash
# One gpurun session (edit per experiment): validation + A/B timings + ncu capture.  Run from the repo root on the GPU box.
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 600 python -m pytest tests/test_gpu_unfused_api.py -x -q -m gpu
export MNF_GATHER_IMPL=4
run timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gather or pack or render"
run timeout 600 python -m pytest tests/test_gpu_unfused_api.py tests/test_gpu_model.py -q -m gpu
run timeout 300 python tools/prof_kernels.py --which gather --reps 5
run timeout 300 python tools/prof_kernels.py --which gather --reps 5 --samples 128 --rays 40960
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_mma_kernel -c 1 -f -o gpurun_out/gather_v4 python tools/prof_kernels.py --which gather --rays 40960 --reps 1
run timeout 600 python bench.py --steps 10 --warmup 3
unset MNF_GATHER_IMPL
run timeout 300 python tools/prof_kernels.py --which gather --reps 5
tail -5 $L
