#!/bin/bash
# Build decoder-kernel variants from git history side by side (A/B on one GPU box).  Usage: [NVCC_EXTRA=-DMNF_DECODER_TRACE] tools/build_variants.sh tag=commit|WORK ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/matchnerf_b200/variants
mkdir -p $OUT
for spec in "$@"; do
  tag=${spec%%=*}; rev=${spec#*=}
  W=/tmp/variant_$tag; rm -rf $W; mkdir -p $W/matchnerf_b200/csrc $W/include
  cp $ROOT/include/*.h $W/include/; cp $ROOT/matchnerf_b200/csrc/*.cu $ROOT/matchnerf_b200/csrc/*.cuh $W/matchnerf_b200/csrc/
  if [ "$rev" != "WORK" ]; then git -C $ROOT show $rev:matchnerf_b200/csrc/decoder_tc.cu > $W/matchnerf_b200/csrc/decoder_tc.cu; fi
  (cd $W/matchnerf_b200/csrc && for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $NVCC_EXTRA -c $f -o ${f%.cu}.o & done; wait; nvcc -shared -o $OUT/lib_$tag.so *.o -gencode arch=compute_100a,code=sm_100a)
  echo built $OUT/lib_$tag.so
done
