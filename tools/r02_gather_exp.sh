#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_gather_exp}.log
(timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "gather or render or model or slicing" 2>&1 | tail -5
 python tools/prof_kernels.py --rays 327680 --which gather --reps 5
 python tools/prof_kernels.py --rays 327680 --samples 128 --which gather --reps 3) > $L 2>&1
tail -4 $L | cut -c1-200
