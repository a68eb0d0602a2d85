#!/bin/bash
# multi-GPU evidence (one box, N GPUs): 2-GPU NCCL tests, strong scaling (ONE image per step sharded over N ranks) at N = 2, 4, 8,
# weak scaling (one image per rank) at N = 8
mkdir -p gpurun_out
L=gpurun_out/r02_multi_s3.log
: > $L
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG" >> $L
(timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -3) >> $L 2>&1
run() { N=$1; SH=$2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 3 --shard $SH --quick 2>&1 | grep '^{"metric"' >> $L; }
python bench.py --steps 20 --warmup 3 --quick 2>&1 | grep '^{"metric"' >> $L
for N in 2 4 8; do if [ $N -le $NG ]; then run $N rays; fi; done
if [ 8 -le $NG ]; then run 8 images; elif [ 4 -le $NG ]; then run 4 images; fi
python - <<PY >> $L
import json
for ln in open("$L"):
    if ln.startswith('{"metric"'):
        d = json.loads(ln)
        print(d["n_gpus"], d["scaling"], "ms/step %.3f" % d["ms_per_step"], "value %.2f M rays/s" % (d["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6))
PY
tail -8 $L
