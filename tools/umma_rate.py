#!/usr/bin/env python
"""[needs a library built with -DMNF_MICROBENCH:  NVCC_EXTRA=-DMNF_MICROBENCH tools/build_variants.sh bench=WORK; MNF_LIB_PATH=matchnerf_b200/variants/lib_bench.so python tools/umma_rate.py]
tcgen05.mma rate micro-benchmark (one CTA on one SM): cycles per M=128 x N x K=16 step, A from shared memory (SS) or
from tensor memory (TS), with 0..4 warps streaming tcgen05.ld meanwhile."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matchnerf_b200 import capi
lib = capi.load()
lib.mnf_selftest_umma_rate.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
lib.mnf_selftest_umma_rate.restype = C.c_int32
out = torch.zeros(4, dtype=torch.int64, device="cuda")
iters = 64
for mode in (0, 1):
    for N, n_acc in ((64, 1), (64, 2), (64, 4), (128, 1), (128, 2), (256, 1)):
        for readers in (0, 4):
            out.zero_()
            rc = lib.mnf_selftest_umma_rate(iters, N, mode, readers, n_acc, out.data_ptr(), None)
            assert rc == 0, capi.load().mnf_last_error()
            torch.cuda.synchronize()
            tot, iss, nrd = int(out[0]), int(out[1]), int(out[2])
            print(f"{'SS' if mode == 0 else 'TS'} N={N:3d} accumulators={n_acc} readers={readers}: {tot / (iters * 8):6.1f} cycles per K=16 step "
                  f"(issue loop {iss / (iters * 8):5.1f}); reader warp 4 streamed {nrd} x 128 columns", flush=True)
