#!/bin/bash
# One gpurun session (edit per experiment): validation + A/B timings + ncu capture.  Run from the repo root on the GPU box.
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 600 python bench.py
run timeout 600 python bench.py --impl reference --steps 2 --warmup 1
run timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -c 1 -f -o gpurun_out/decoder_r01_final python tools/prof_kernels.py --which decoder --impl 2 --rays 40960 --reps 1
tail -5 $L
