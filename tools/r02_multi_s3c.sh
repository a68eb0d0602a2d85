#!/bin/bash
# 8-GPU refresh on the session-3 state: strong (one image sharded over 8 ranks) and weak (one image per rank), quick lines
mkdir -p gpurun_out
L=gpurun_out/r02_multi_s3c.log
: > $L
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG" >> $L
run() { N=$1; SH=$2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 3 --shard $SH --quick 2>&1 | grep '^{"metric"' >> $L; }
run $NG rays
run $NG images
if [ $NG -ge 8 ]; then run 4 rays; fi
python - <<PY >> $L
import json
for ln in open("$L"):
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        print(d["n_gpus"], d["scaling"], "ms/step %.3f" % d["ms_per_step"], "value %.2f M rays/s" % (d["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6))
PY
tail -4 $L
