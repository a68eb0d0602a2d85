#!/bin/bash
# One gpurun session: gather v3 validation + A/B, encoder variants, full GPU tests, bench, ncu capture.
# Usage (from the repo root on the GPU box): bash tools/gpu_call.sh
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
run timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gather or pack or render"
MNF_GATHER_IMPL=2 run timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gather or pack"
echo "=== A/B gather (v3 mixed / v3 ffma2 / v2)" >> $L
run timeout 300 python tools/prof_kernels.py --which gather --reps 5
MNF_GATHER_MIXED=0 run timeout 300 python tools/prof_kernels.py --which gather --reps 5
MNF_GATHER_IMPL=2 run timeout 300 python tools/prof_kernels.py --which gather --reps 5
run timeout 300 python tools/prof_kernels.py --which gather --reps 5 --samples 128 --rays 40960
run timeout 300 python tools/prof_encoder.py
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_cossim_kernel -c 1 -f -o gpurun_out/gather_v3 python tools/prof_kernels.py --which gather --rays 40960 --reps 1
run timeout 1200 python -m pytest tests -x -q -m gpu
run timeout 600 python bench.py
tail -5 $L
