#!/bin/bash
# One gpurun session: full validation + bench of both arms.  Run from the repo root on the GPU box:  bash tools/gpu_call.sh
set -u
mkdir -p gpurun_out
L=gpurun_out/call.log
: > $L
run() { echo "=== $*" >> $L; ( "$@" ) >> $L 2>&1; echo "--- exit $?" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu -x
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 600 python bench.py
run timeout 600 python bench.py --impl reference --steps 2 --warmup 1
tail -5 $L
