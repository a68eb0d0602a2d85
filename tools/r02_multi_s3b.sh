#!/bin/bash
# the driver's own N = 2 commands: default bench (weak, image-parallel) and the reference arm under torchrun
mkdir -p gpurun_out
L=gpurun_out/r02_multi_s3b.log
: > $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -4 >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3 >> $L
cut -c1-700 $L
