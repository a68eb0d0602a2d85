#!/bin/bash
mkdir -p gpurun_out
MNF_LIB_PATH=matchnerf_b200/variants/lib_trace.so python tools/decoder_trace.py ${2:-} > gpurun_out/${1:-r02_trace}.log 2>&1
tail -3 gpurun_out/${1:-r02_trace}.log
