#!/bin/bash
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_cossim_kernel -c 1 -f -o gpurun_out/gather_v3_final python tools/prof_kernels.py --which gather --rays 40960 --reps 1
tail -5 $L
