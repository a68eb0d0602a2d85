#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_train_exp}.log
(timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -30) > $L 2>&1
tail -30 $L | cut -c1-250
