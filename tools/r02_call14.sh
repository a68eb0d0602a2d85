#!/bin/bash
# session-3 state: full GPU suite (token path, metrics, fused projection, local radius), bench, new-kernel ncu captures
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call14}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu
run timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train-step --no-reference-gpu
run timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_enc_launches2.csv python tools/enc_launches.py
run timeout 300 ncu --set full --clock-control none --import-source on -k regex:token_block_kernel -s 12 -c 1 -f -o gpurun_out/r02_token_block python tools/time_token_block.py --reps 2
MNF_TOKEN_PATH=0 run timeout 600 python bench.py --steps 10 --warmup 3 --quick
grep -n "^===\|^--- exit\|passed\|failed\|Error\|error\|assert" $L | cut -c1-260 | head -60
