#!/bin/bash
# One gpurun session (edit per experiment): validation + A/B timings + ncu capture.  Run from the repo root on the GPU box.
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "window_attn" -x
run timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -x
run timeout 300 python tools/prof_kernels.py --which attn --reps 10 --impl 2
MNF_ATTN_PIPE=1 run timeout 300 python tools/prof_kernels.py --which attn --reps 10 --impl 2
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_tc_v4_kernel -c 1 -f -o gpurun_out/attn_v4 python tools/prof_kernels.py --which attn --impl 2 --reps 1
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attn" -c 8 python tools/prof_kernels.py --which attn --impl 2 --reps 1
run timeout 300 python tools/prof_encoder.py
tail -5 $L
