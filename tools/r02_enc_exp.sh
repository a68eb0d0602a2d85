#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_enc_exp}.log
(timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
 python tools/prof_kernels.py --rays 40960 --which encoder --reps 5
 timeout 300 python bench.py --steps 10 --warmup 3 --quick | tail -1 | cut -c1-400) > $L 2>&1
tail -14 $L | cut -c1-300
