#!/bin/bash
# Round-2 evidence call: both bench arms with default flags (wall-clocked), ncu launch list of the bench command, one
# `ncu --set full` capture each of the tensor-core gather and the decoder, compute-sanitizer memcheck / racecheck passes.
set -u
mkdir -p gpurun_out
L=gpurun_out/r02_call8.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
python -c "import __graft_entry__ as g; g.build()" >> $L 2>&1
run timeout 900 python bench.py --impl reference
run timeout 900 python bench.py
grep '^{"metric"' $L | tail -1 > gpurun_out/r02_bench_call8.json
grep '^{"impl"' $L | tail -1 > gpurun_out/r02_bench_call8_reference_arm.json
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel -c 1 -f -o gpurun_out/r02_decoder \
    python tools/prof_kernels.py --rays 40960 --which decoder --impl 2 --reps 1
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_tc -c 1 -f -o gpurun_out/r02_gather_tc_v6 \
    python tools/prof_kernels.py --rays 40960 --which gather --reps 1
# sanitizer passes over the kernel tests (bounded: small cases only)
run timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
    python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "borders_ragged or small_golden or window_attn_golden or instance_norm or edge_cases or tensor_core_path"
run timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 \
    python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "borders_ragged or tcgen05_small_golden or window_attn_golden"
tail -5 $L | cut -c1-300
