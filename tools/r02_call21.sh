#!/bin/bash
# final validation of round 2: full GPU suite, smoke, default bench (incl. the training step in fp32 and TF32)
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_call21}.log
: > $L
run() { echo "=== $*" >> $L; local t0=$SECONDS; ( "$@" ) >> $L 2>&1; echo "--- exit $? after $((SECONDS - t0)) s" >> $L; }
run timeout 900 python -m pytest tests -q -m gpu
run timeout 600 python -c "import __graft_entry__ as g; g.smoke()"
run timeout 900 python bench.py --no-cpu-baseline
grep -n "^===\|^--- exit\|passed\|failed\|smoke:\|Error" $L | cut -c1-220
python - <<PY
import json
for ln in open("$L"):
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"]); print("train", {k: v for k, v in d["train_step"].items() if not isinstance(v, str)})
PY
