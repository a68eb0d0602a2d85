#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "window_attn" -x
run timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "encoder or forward"
run timeout 300 python tools/prof_kernels.py --which attn --reps 10 --impl 2
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn" -c 4 python tools/prof_kernels.py --which attn --impl 2 --reps 1
tail -5 $L
