#!/bin/bash
# fused projection bring-up + encoder launch list + token_block timing
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call13}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused_projection or token_block or window_attn"
run timeout 120 python tools/time_token_block.py
run timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin_demo.py -q -m gpu
run timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-step --no-reference-gpu
run timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_enc_launches.csv python tools/enc_launches.py
MNF_FUSED_PROJ=0 run timeout 600 python bench.py --steps 10 --warmup 3 --quick
grep -n "^===\|^--- exit\|passed\|failed\|token_block rows\|Error\|error\|assert" $L | cut -c1-260 | head -60
