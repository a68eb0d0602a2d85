#!/bin/bash
# time decoder build variants (matchnerf_b200/variants/lib_<tag>.so) side by side: tools/r02_variants.sh <log> <tag> ...
mkdir -p gpurun_out
L=gpurun_out/$1.log; shift
: > $L
for tag in "$@"; do
  echo "== variant $tag" >> $L
  for S in 64 128; do
    if [ "$tag" = base ]; then python tools/prof_kernels.py --rays 327680 --samples $S --which decoder --impl 2 --reps 5 2>&1 | grep decoder >> $L
    else MNF_LIB_PATH=matchnerf_b200/variants/lib_$tag.so python tools/prof_kernels.py --rays 327680 --samples $S --which decoder --impl 2 --reps 5 2>&1 | grep decoder >> $L; fi
  done
done
cat $L
