#!/bin/bash
# decoder experiment: parity tests of the decoder / render paths, isolated timing at S = 64 / 128 / 256, timeline of CTA 0
set -u
mkdir -p gpurun_out
L=gpurun_out/${1:-r02_dec_exp}.log
: > $L
(timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -m gpu -x -k "decoder or render or model or slicing or video" 2>&1 | tail -5
 python tools/prof_kernels.py --rays 327680 --which decoder --impl 2 --reps 5
 python tools/prof_kernels.py --rays 327680 --samples 128 --which decoder --impl 2 --reps 3
 python tools/prof_kernels.py --rays 81920 --samples 256 --which decoder --impl 2 --reps 3
 if [ "${2:-}" = "trace" ]; then
   MNF_LIB_PATH=matchnerf_b200/variants/lib_trace.so python tools/decoder_trace.py --samples ${3:-64}
 fi) >> $L 2>&1
tail -4 $L | cut -c1-200
