#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/r02_call10.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "borders_ragged or tcgen05_small_golden or tensor_core_path"
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_tc -c 1 -f -o gpurun_out/r02_gather_tc_v7 python tools/prof_kernels.py --rays 40960 --which gather --reps 1
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -c 1 -f -o gpurun_out/r02_decoder_final python tools/prof_kernels.py --rays 40960 --which decoder --impl 2 --reps 1
grep -n "^===\|^--- exit\|passed\|failed\|SUMMARY" $L | cut -c1-160
