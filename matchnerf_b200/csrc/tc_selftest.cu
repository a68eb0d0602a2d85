// Self-test of the tcgen05 building blocks the decoder / attention kernels are made of.
// D[128][N] = A[128][K] * B[N][K]^T, fp16 operands, fp32 accumulation in tensor memory, one CTA of 128 threads.
//   mode 0 (SS): A and B in shared memory, SWIZZLE_128B K-major tiles of 64 columns, written with plain stores.
//   mode 1 (TS): A written to tensor memory with tcgen05.st (two fp16 per column), B in shared memory.
//   mode 2 (TS, MN-major B): as mode 1, but B is given as [K][N] (N contiguous) and used as an MN-major operand --
//           the layout of the V tile in attention (rows = keys, contiguous value channels).
// tests/test_gpu_tc.py compares the result with a float64 product of the same fp16 inputs.
#include <cuda_fp16.h>

#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {
constexpr int kTmemCols = 512;
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, const int N,
                     const int K, const int mode) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  // SWIZZLE_128B tiles need a 1024 B aligned base; do not rely on the toolchain for it
  unsigned char* smem = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kblocks = K / 64;
  // smem carve-up: A tiles [kblocks][128 x 128 B] (mode 0 only), then B tiles [kblocks][N x 128 B]; all 1024 B multiples
  unsigned char* sA = smem;
  unsigned char* sB = smem + (mode == 0 ? (size_t)kblocks * 128 * 128 : 0);

  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<kTmemCols>(&tmem_base_slot);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t d_tmem = tmem;            // columns [0, N)
  const uint32_t a_tmem = tmem + 256;      // columns [256, 256 + K/2) in TS mode

  // ---- stage B (and A in SS mode) into the swizzled layout, 16 B chunks
  if (mode != 2) {
    for (int i = tid; i < N * (K / 8); i += blockDim.x) {
      const int row = i / (K / 8), ch = i - row * (K / 8);
      const int kb = ch / 8, c8 = ch & 7;
      const uint4 val = *reinterpret_cast<const uint4*>(B + (size_t)row * K + ch * 8);
      *reinterpret_cast<uint4*>(sB + (size_t)kb * N * 128 + tc::sw128_offset(row, c8 * 8)) = val;
    }
  } else {
    // B[K][N]: tile nb holds columns [64 nb, 64 nb + 64) of all K rows
    for (int i = tid; i < K * (N / 8); i += blockDim.x) {
      const int row = i / (N / 8), ch = i - row * (N / 8);
      const int nb = ch / 8, c8 = ch & 7;
      const uint4 val = *reinterpret_cast<const uint4*>(B + (size_t)row * N + ch * 8);
      *reinterpret_cast<uint4*>(sB + (size_t)nb * K * 128 + tc::sw128_offset(row, c8 * 8)) = val;
    }
  }
  if (mode == 0) {
    for (int i = tid; i < 128 * (K / 8); i += blockDim.x) {
      const int row = i / (K / 8), ch = i - row * (K / 8);
      const int kb = ch / 8, c8 = ch & 7;
      const uint4 val = *reinterpret_cast<const uint4*>(A + (size_t)row * K + ch * 8);
      *reinterpret_cast<uint4*>(sA + (size_t)kb * 128 * 128 + tc::sw128_offset(row, c8 * 8)) = val;
    }
  } else {
    // thread = row; 16 columns (32 halves) per tcgen05.st
    const uint32_t lane_base = (uint32_t)(warp * 32);
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + (c0 + j) * 2);
      tc::tmem_st16(tc::tmem_addr(a_tmem, lane_base, c0), r);
    }
    tc::tmem_wait_st();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();

  if (tid == 0) {
    tc::tc_fence_after_sync();
    const uint32_t idesc = tc::umma_idesc_f16(128, N, mode == 2 ? 1u : 0u);
    if (mode == 2) {
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t bdesc = tc::umma_desc_sw128_mn(tc::smem_u32(sB) + ks * 2048, (uint32_t)K * 128u);
        tc::umma_ts(d_tmem, a_tmem + ks * 8, bdesc, idesc, ks ? 1u : 0u);
      }
    }
    for (int kb = 0; kb < (mode == 2 ? 0 : kblocks); ++kb) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // 4 x K=16 per 64-wide block; +32 B per step inside the swizzle atom
        const uint64_t bdesc = tc::umma_desc_sw128(tc::smem_u32(sB + (size_t)kb * N * 128) + ks * 32);
        const uint32_t acc = (kb | ks) ? 1u : 0u;
        if (mode == 0) {
          const uint64_t adesc = tc::umma_desc_sw128(tc::smem_u32(sA + (size_t)kb * 128 * 128) + ks * 32);
          tc::umma_ss(d_tmem, adesc, bdesc, idesc, acc);
        } else {
          tc::umma_ts(d_tmem, a_tmem + (kb * 64 + ks * 16) / 2, bdesc, idesc, acc);
        }
      }
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after_sync();

  // ---- read the accumulator back: thread = row, 32 columns per load
  const uint32_t lane_base = (uint32_t)(warp * 32);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tc::tmem_ld32(tc::tmem_addr(d_tmem, lane_base, c0), r);
    tc::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace mnf

extern "C" int32_t mnf_selftest_umma(const void* a_f16, const void* b_f16, float* d_f32, int32_t N, int32_t K, int32_t mode,
                                     void* stream) {
  using namespace mnf;
  if (!a_f16 || !b_f16 || !d_f32 || K % 64 != 0 || K <= 0 || K > 256 || N % 16 != 0 || N < 16 || N > 256 || (mode < 0 || mode > 2) || (mode == 2 && N % 64 != 0)) {
    set_error("mnf_selftest_umma: bad arguments (N=%d K=%d mode=%d)", N, K, mode);
    return MNF_EINVAL;
  }
  const size_t smem = (size_t)(K / 64) * 128 * 128 + (size_t)N * K * 2 + 1024;
  MNF_CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(a_f16),
                                                              reinterpret_cast<const __half*>(b_f16), d_f32, N, K, mode);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

#ifdef MNF_MICROBENCH   // tensor-memory / UMMA rate micro-benchmarks (tools/tmem_bw.py, tools/umma_rate.py): built only with -DMNF_MICROBENCH
// ---- micro-benchmark: tensor-memory read (tcgen05.ld) throughput of one SM ------------------------------------------
// `warps` warps (4 or 8; two warps share a lane quarter when 8) each read `cols` columns `iters` times; out[0] = cycles.
namespace mnf {
__global__ void __launch_bounds__(256, 1) tmem_bw_kernel(int iters, int cols, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 4 ? 256 : 0);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    for (int c0 = 0; c0 < cols; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tb + c0, r);
      tc::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; }
  if (acc == 0x12345678u) out[1] = acc;   // keep the loads alive
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(slot);
}
}  // namespace mnf

extern "C" int32_t mnf_selftest_tmem_bw(int32_t warps, int32_t iters, int32_t cols, long long* cycles_out_dev, void* stream) {
  using namespace mnf;
  if ((warps != 4 && warps != 8) || iters <= 0 || cols <= 0 || cols > 256 || cols % 32 || !cycles_out_dev) {
    set_error("mnf_selftest_tmem_bw: bad arguments");
    return MNF_EINVAL;
  }
  tmem_bw_kernel<<<1, warps * 32, 0, (cudaStream_t)stream>>>(iters, cols, cycles_out_dev);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

// ---- micro-benchmark: tcgen05.mma issue / execution rate of one SM --------------------------------------------------
// One CTA.  `iters` back-to-back groups of 8 K=16 steps (M = 128, N columns, fp16, fp32 accumulate) are issued by one
// thread and committed once; out[0] = cycles from the first issue to the commit's arrival, out[1] = cycles the issue loop
// itself took.  mode 0: A and B from shared memory; mode 1: A from tensor memory.  `readers` > 0: that many extra warps
// (1..4, one per TMEM lane quarter) stream tcgen05.ld over 128 accumulator columns of another region meanwhile.
namespace mnf {
__global__ void __launch_bounds__(256, 1) umma_rate_kernel(int iters, int N, int mode, int readers, int n_acc, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 128 + 256 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
    done = 0;
  }
  if (warp == 0) tc::tmem_alloc<512>(&slot);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = slot;
  if (warp < 4) {   // zero the A operand columns [256, 320)
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    for (int c0 = 0; c0 < 64; c0 += 16) tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, z);
    tc::tmem_wait_st();
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::umma_idesc_f16(128, N);
    unsigned char* sA = smem;
    unsigned char* sB = smem + 128 * 128;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t bdesc = tc::umma_desc_sw128(tc::smem_u32(sB) + (ks & 3) * 32);
        // n_acc independent accumulators (column regions of N) used round-robin: consecutive MMAs then do not depend on each other
        const uint32_t d = tmem + (uint32_t)((ks % n_acc) * N);
        if (mode == 0) tc::umma_ss(d, tc::umma_desc_sw128(tc::smem_u32(sA) + (ks & 3) * 32), bdesc, idesc, 1u);
        else tc::umma_ts(d, tmem + 256 + ks * 8, bdesc, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0;
    out[1] = t1 - t0;
    done = 1;
  } else if (warp >= 4 && warp < 4 + readers) {
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 320;
    uint32_t acc = 0;
    long long n = 0;
    while (!done) {
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tb + c0, r);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
      ++n;
    }
    if (acc == 0x12345678u) out[3] = acc;
    if ((tid & 31) == 0 && warp == 4) out[2] = n;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}
}  // namespace mnf

extern "C" int32_t mnf_selftest_umma_rate(int32_t iters, int32_t N, int32_t mode, int32_t readers, int32_t n_acc, long long* out_dev, void* stream) {
  using namespace mnf;
  if (iters <= 0 || N < 16 || N > 256 || N % 16 || mode < 0 || mode > 1 || readers < 0 || readers > 4 || !out_dev || n_acc < 1 || n_acc > 4 ||
      n_acc * N > 256) {
    set_error("mnf_selftest_umma_rate: bad arguments");
    return MNF_EINVAL;
  }
  const size_t smem = 128 * 128 + 256 * 128 + 1024;
  MNF_CUDA_TRY(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(iters, N, mode, readers, n_acc, out_dev);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}
#endif  // MNF_MICROBENCH
