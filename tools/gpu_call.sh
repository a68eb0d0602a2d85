#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
MNF_GATHER_IMPL=4 run timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "gather or pack or render or forward or train_mode"
MNF_ATTN_PIPE=1 run timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "window_attn"
run timeout 600 python bench.py --no-cpu-baseline --steps 10
tail -5 $L
