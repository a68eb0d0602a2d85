#!/bin/bash
# One gpurun session (edit per experiment): validation + A/B timings + ncu capture.  Run from the repo root on the GPU box.
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "window_attn"
run timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu
run timeout 300 python tools/prof_kernels.py --which attn --reps 10 --impl 2
run timeout 300 python tools/prof_encoder.py
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_tc_kernel -c 2 -f -o gpurun_out/attn_v2 python tools/prof_kernels.py --which attn --impl 2 --reps 1
tail -5 $L
