#!/bin/bash
# Round-2 session-3 evidence call: full GPU suite, smoke, both bench arms (default flags), ncu launch list of the bench command,
# ncu --set full of the new kernels, sanitizer passes (memcheck + racecheck) over the kernels added this session
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call16}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 900 python bench.py --impl reference
run timeout 900 python bench.py
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --quick
run timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_project_pack -s 6 -c 1 -f -o gpurun_out/r02_attn_project_pack python tools/enc_launches.py
run timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_metrics.py -q -m gpu -x -k "token_block or fused_projection or local_radius or instance_norm_nhwc or eval_tools or small_golden"
run timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_metrics.py -q -m gpu -x -k "token_block or fused_projection or local_radius or eval_tools"
grep -n "^===\|^--- exit\|passed\|failed\|SUMMARY\|smoke:" $L | cut -c1-220
