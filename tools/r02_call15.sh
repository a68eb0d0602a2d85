#!/bin/bash
# token_block with 16 epilogue warps + instance-norm statistics tail: parity, timing, bench
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call15}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -k "token_block or instance_norm or encoder or forward_small or slicing"
run timeout 120 python tools/time_token_block.py
run timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-step --no-reference-gpu
run timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_enc_launches3.csv python tools/enc_launches.py
run timeout 300 ncu --set full --clock-control none --import-source on -k regex:token_block_kernel -s 12 -c 1 -f -o gpurun_out/r02_token_block_v2 python tools/time_token_block.py --reps 2
grep -n "^===\|^--- exit\|passed\|failed\|token_block rows\|Error\|error\|assert" $L | cut -c1-260 | head -60
