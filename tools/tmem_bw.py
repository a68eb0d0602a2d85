#!/usr/bin/env python
"""[needs a library built with -DMNF_MICROBENCH:  NVCC_EXTRA=-DMNF_MICROBENCH tools/build_variants.sh bench=WORK; MNF_LIB_PATH=matchnerf_b200/variants/lib_bench.so python tools/tmem_bw.py]
Tensor-memory read throughput micro-benchmark (tcgen05.ld 32x32b.x32), one CTA on one SM."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matchnerf_b200 import capi
lib = capi.load()
lib.mnf_selftest_tmem_bw.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
lib.mnf_selftest_tmem_bw.restype = C.c_int32
out = torch.zeros(2, dtype=torch.int64, device="cuda")
for warps in (4, 8):
    for cols in (128, 256):
        iters = 2000
        assert lib.mnf_selftest_tmem_bw(warps, iters, cols, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        cyc = int(out[0])
        byts = warps * 32 * cols * 4 * iters
        print(f"warps={warps} cols={cols}: {cyc} cycles for {byts/1e6:.1f} MB -> {byts/cyc:.1f} B/cycle/SM, {cyc/iters/(cols//32):.1f} cycles per ld32+wait per warp")
