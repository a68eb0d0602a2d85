#!/bin/bash
# Round-2 validation call: full GPU suite, smoke, both bench arms (default flags, wall-clocked), sanitizer passes over the new kernels
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call9}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 900 python bench.py --impl reference
run timeout 900 python bench.py
run timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -q -m gpu -x -k "borders_ragged or small_golden or window_attn_golden or instance_norm or token_layernorm or edge_cases or tensor_core_path or gather_backward"
run timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "borders_ragged or tcgen05_small_golden or tensor_core_path"
grep -n "^===\|^--- exit\|passed\|failed\|SUMMARY" $L | cut -c1-200
