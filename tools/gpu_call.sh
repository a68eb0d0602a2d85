#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 300 python tools/attn_trace.py
run timeout 300 python tools/attn_trace.py --shift
run timeout 300 python tools/prof_kernels.py --which attn --reps 10 --impl 2
tail -5 $L
