// K-block: everything of a GMFlow TransformerLayer that follows the attention, for a 128-token tile, in ONE tcgen05 kernel.
//
// Replaces TransformerLayer.forward after the attention call (models/gmflow/transformer.py:173-185):
//     message = merge(message)                    Linear 128 -> 128, no bias
//     message = norm1(message)                    LayerNorm(128)
//     if not no_ffn:
//         message = mlp(cat[source, message])     Linear 256 -> 1024 (no bias), GELU (erf form), Linear 1024 -> 128 (no bias)
//         message = norm2(message)
//     return source + message
// Before: cuBLAS TF32 merge GEMM -> LayerNorm kernel (+ cat + fp16) -> cuBLAS fp16 GEMM -> ATen GELU -> cuBLAS fp16 GEMM ->
// LayerNorm kernel (+ residual): six launches per cross-attention layer and a 63 MB round trip of the 1024-wide hidden tensor
// through L2 / HBM.  Here the hidden tensor never leaves the SM: per 128-column hidden chunk j,
//     acc1[j&1] = X (128 x 256, shared memory) . W1_j^T          tcgen05.mma SS, 16 K-steps
//     H[j&1]    = fp16(GELU(acc1[j&1]))                         epilogue warps: tcgen05.ld -> math -> tcgen05.st (A operand in TMEM)
//     acc2     += H[j&1] (TMEM) . W2_j^T                          tcgen05.mma TS, 8 K-steps
// with the weights streamed through an 8-stage shared-memory ring of pre-swizzled 16 KB tiles (cp.async.bulk + mbarrier tx).
//
// One persistent CTA per SM, 640 threads:
//   warp 0       weight streamer (one lane)
//   warp 1       MMA issuer (one elected lane), owns the TMEM allocation
//   warps 4-19   16 epilogue warps: thread = token row = TMEM lane (warp % 4 = lane quarter), four warps per quarter split the
//                128 columns of an accumulator in 32-column quarters.  They also load + convert the tile's inputs (prologue), evaluate both
//                LayerNorms (a thread sees its whole row in TMEM: no cross-thread reduction) and write the result.
// TMEM (512 columns): acc1[0] [0,128)  acc1[1] [128,256)  acc2 [256,384)  H[0] [384,448)  H[1] [448,512).
// Shared memory: X = 4 K-blocks [128 rows][64 fp16] SWIZZLE_128B (K-blocks 0,1 = source, 2,3 = LN1 output; before the merge
// product K-blocks 2,3 hold the attention tile, its A operand), ring 8 x 16 KB.
//
// Arithmetic: fp16 operands (10-bit mantissa, the same the TF32 GEMMs this replaces rounded to), fp32 accumulation, LayerNorm /
// GELU / residual in fp32.  GELU(x) = max(x, 0) - |x| q(|x|), q(t) = 0.5 erfc(t / sqrt 2) = 2^(-1 - t P(t)) with a degree-4
// polynomial P fitted on [0, 6]: max abs error 7.7e-7 against the erf form (the hidden value is then rounded to fp16, 2.4e-4).
#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {

constexpr int kTok = 128;                    // tokens per tile = MMA M
constexpr int kCh = 128;                     // d_model
constexpr int kHid = 1024;                   // ffn hidden width (2 * d_model * ffn_dim_expansion)
constexpr int kHidChunks = kHid / 128;       // 8
constexpr int kTileBytes = 128 * 128;        // one [128 rows][64 fp16] SWIZZLE_128B tile
constexpr int kStages = 8;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (4 + kEpiWarps) * 32;      // 640
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr uint32_t kColAcc1 = 0, kColAcc2 = 256, kColH = 384;
constexpr int kTilesFfn = 2 + 32 + 16, kTilesNoFfn = 2;

struct BlockBars {
  uint64_t full[kStages], empty[kStages];
  uint64_t a_ready, x_ready, acc2_full;
  uint64_t acc1_full[2], acc1_empty[2], h_ready[2], h_empty[2];
  uint32_t tmem_base;
};

constexpr size_t kSmemBytes = 1024 + 4 * kTileBytes + (size_t)kStages * kTileBytes + sizeof(BlockBars);

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// GELU, erf form (nn.GELU default), see the header comment
__device__ __forceinline__ float gelu_erf(float x) {
  const float t = fminf(fabsf(x), 6.0f);
  float p = 0.000494403182528913f;
  p = fmaf(p, t, -0.007237763609737158f);
  p = fmaf(p, t, 0.05222909152507782f);
  p = fmaf(p, t, 0.4595268964767456f);
  p = fmaf(p, t, 1.1510192155838013f);
  float q;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(fmaf(-t, p, -1.0f)));      // exponent in [-31, -1]: no range handling needed
  return fmaxf(x, 0.0f) - t * q;
}

// mean and 1/sqrt(var + eps) of this thread's row of a 128-column fp32 accumulator (two passes over TMEM: centred variance)
__device__ __forceinline__ void row_stats(uint32_t acc_lane, float eps, float& mean, float& rstd) {
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kCh; c += 32) {
    uint32_t r[32];
    tc::tmem_ld32(acc_lane + c, r);
    tc::tmem_wait_ld(r);
#pragma unroll
    for (int i = 0; i < 32; ++i) sum += __uint_as_float(r[i]);
  }
  mean = sum * (1.0f / kCh);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < kCh; c += 32) {
    uint32_t r[32];
    tc::tmem_ld32(acc_lane + c, r);
    tc::tmem_wait_ld(r);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = __uint_as_float(r[i]) - mean;
      ss = fmaf(d, d, ss);
    }
  }
  rstd = rsqrtf(ss * (1.0f / kCh) + eps);
}

__global__ void __launch_bounds__(kThreads, 1)
token_block_kernel(const float* __restrict__ attn, const float* __restrict__ source, const unsigned char* __restrict__ wpk,
                   float* __restrict__ out, const int64_t T, const int with_ffn, const float eps) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sX = smem;
  unsigned char* sRing = smem + 4 * kTileBytes;
  BlockBars& sm = *reinterpret_cast<BlockBars*>(sRing + (size_t)kStages * kTileBytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (T + kTok - 1) / kTok;
  const float* ln = reinterpret_cast<const float*>(wpk + (size_t)(with_ffn ? kTilesFfn : kTilesNoFfn) * kTileBytes);   // g1 b1 g2 b2

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      tc::mbar_init(&sm.full[i], 1);
      tc::mbar_init(&sm.empty[i], 1);
    }
    tc::mbar_init(&sm.a_ready, kEpiThreads);
    tc::mbar_init(&sm.x_ready, kEpiThreads);
    tc::mbar_init(&sm.acc2_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sm.acc1_full[i], 1);
      tc::mbar_init(&sm.acc1_empty[i], kEpiThreads);
      tc::mbar_init(&sm.h_ready[i], kEpiThreads);
      tc::mbar_init(&sm.h_empty[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ================================================================== weight streamer: tiles in the order the MMA warp reads them
    if (lane == 0) {
      uint32_t n = 0;
      auto push = [&](int tile_idx) {
        const uint32_t st = n % kStages, par = (n / kStages) & 1u;
        tc::mbar_wait_sleep(&sm.empty[st], par ^ 1u, 64);
        tc::mbar_arrive_expect_tx(&sm.full[st], kTileBytes);
        tc::bulk_g2s(sRing + (size_t)st * kTileBytes, wpk + (size_t)tile_idx * kTileBytes, kTileBytes, &sm.full[st]);
        ++n;
      };
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        push(0);
        push(1);
        if (with_ffn) {
          for (int j = 0; j <= kHidChunks; ++j) {
            if (j < kHidChunks)
              for (int kb = 0; kb < 4; ++kb) push(2 + j * 4 + kb);
            if (j > 0)
              for (int kb = 0; kb < 2; ++kb) push(34 + (j - 1) * 2 + kb);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::umma_idesc_f16(128, 128);
    const uint32_t x_addr = tc::smem_u32(sX), ring_addr = tc::smem_u32(sRing);
    uint32_t n = 0, G = 0, C = 0, it = 0;
    auto next_stage = [&]() {
      const uint32_t st = n % kStages, par = (n / kStages) & 1u;
      tc::mbar_wait(&sm.full[st], par);
      ++n;
      return st;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      tc::mbar_wait(&sm.a_ready, it & 1u);
      tc::tc_fence_after_sync();
      {   // merge: acc1[b] = attention tile (X K-blocks 2, 3) . Wm^T
        const uint32_t b = G & 1u;
        tc::mbar_wait(&sm.acc1_empty[b], ((G >> 1) & 1u) ^ 1u);
        tc::tc_fence_after_sync();
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t st = next_stage();
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::umma_ss(tmem + kColAcc1 + b * 128, tc::umma_desc_sw128(x_addr + (2 + kb) * kTileBytes + ks * 32),
                          tc::umma_desc_sw128(ring_addr + st * kTileBytes + ks * 32), idesc, (kb | ks) ? 1u : 0u);
            tc::umma_commit(&sm.empty[st]);
          }
          __syncwarp();
        }
        if (leader) tc::umma_commit(&sm.acc1_full[b]);
        __syncwarp();
        ++G;
      }
      if (!with_ffn) continue;
      tc::mbar_wait(&sm.x_ready, it & 1u);
      tc::tc_fence_after_sync();
      for (int j = 0; j <= kHidChunks; ++j) {
        if (j < kHidChunks) {   // FFN1 chunk j: acc1[b] = X . W1_j^T
          const uint32_t b = G & 1u;
          tc::mbar_wait(&sm.acc1_empty[b], ((G >> 1) & 1u) ^ 1u);
          tc::tc_fence_after_sync();
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t st = next_stage();
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc::umma_ss(tmem + kColAcc1 + b * 128, tc::umma_desc_sw128(x_addr + kb * kTileBytes + ks * 32),
                            tc::umma_desc_sw128(ring_addr + st * kTileBytes + ks * 32), idesc, (kb | ks) ? 1u : 0u);
              tc::umma_commit(&sm.empty[st]);
            }
            __syncwarp();
          }
          if (leader) tc::umma_commit(&sm.acc1_full[b]);
          __syncwarp();
          ++G;
        }
        if (j > 0) {            // FFN2 chunk j-1: acc2 += H[hb] . W2_{j-1}^T
          const uint32_t hb = C & 1u;
          tc::mbar_wait(&sm.h_ready[hb], (C >> 1) & 1u);
          tc::tc_fence_after_sync();
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t st = next_stage();
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc::umma_ts(tmem + kColAcc2, tmem + kColH + hb * 64 + (kb * 4 + ks) * 8,
                            tc::umma_desc_sw128(ring_addr + st * kTileBytes + ks * 32), idesc, (j > 1 || (kb | ks)) ? 1u : 0u);
              tc::umma_commit(&sm.empty[st]);
            }
            __syncwarp();
          }
          if (leader) tc::umma_commit(&sm.h_empty[hb]);
          __syncwarp();
          ++C;
        }
      }
      if (leader) tc::umma_commit(&sm.acc2_full);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================================================== epilogue warps
    // ncu of the first version (8 warps, 64 columns per thread): tensor pipe 27 %, issue slots 34 %, 3 long-scoreboard stalls per
    // issue -- two warps per scheduler cannot hide the tcgen05.ld round trips and the MUFU / FMA chains of the GELU, the epilogue
    // (not the MMAs) set the pace.  Now 16 warps: four per lane quarter, 32 columns each, both loads of a chunk in flight at once.
    const int ew = warp - 4, q = warp & 3, cq = ew >> 2;       // cq: which 32 of the 128 columns
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t G = 0, C = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int64_t tok0 = tile * kTok;
      // ---- prologue: attention tile -> X K-blocks 2,3; source tile -> X K-blocks 0,1 (fp16, swizzled).  A warp reads whole rows.
#pragma unroll
      for (int rr = 0; rr < kTok / kEpiWarps; rr += 4) {
        float4 a[4], s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t g = tok0 + ew * (kTok / kEpiWarps) + rr + u;
          if (g < T) {
            a[u] = __ldg(reinterpret_cast<const float4*>(attn + g * kCh) + lane);
            s[u] = __ldg(reinterpret_cast<const float4*>(source + g * kCh) + lane);
          } else {
            a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            s[u] = a[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t r = ew * (kTok / kEpiWarps) + rr + u;
          const uint32_t off = (uint32_t)(lane >> 4) * kTileBytes + tc::sw128_offset(r, (lane & 15) * 4);
          *reinterpret_cast<uint2*>(sX + 2 * kTileBytes + off) = make_uint2(pack_h2(a[u].x, a[u].y), pack_h2(a[u].z, a[u].w));
          *reinterpret_cast<uint2*>(sX + off) = make_uint2(pack_h2(s[u].x, s[u].y), pack_h2(s[u].z, s[u].w));
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&sm.a_ready);

      const int64_t g = tok0 + row;
      const bool valid = g < T;
      const float* src_row = source + (valid ? g : 0) * kCh + cq * 32;
      float* out_row = out + (valid ? g : 0) * kCh + cq * 32;
      // ---- LayerNorm 1 on the merge product
      {
        const uint32_t b = G & 1u;
        tc::mbar_wait(&sm.acc1_full[b], (G >> 1) & 1u);
        tc::tc_fence_after_sync();
        ++G;
        const uint32_t acc = lane_addr + kColAcc1 + b * 128;
        float mean, rstd;
        row_stats(acc, eps, mean, rstd);
        uint32_t r[32];
        tc::tmem_ld32(acc + cq * 32, r);
        tc::tmem_wait_ld(r);
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int col = cq * 32 + i;
          const float4 gm = __ldg(reinterpret_cast<const float4*>(ln + col));
          const float4 bt = __ldg(reinterpret_cast<const float4*>(ln + 128 + col));
          y[i + 0] = fmaf((__uint_as_float(r[i + 0]) - mean) * rstd, gm.x, bt.x);
          y[i + 1] = fmaf((__uint_as_float(r[i + 1]) - mean) * rstd, gm.y, bt.y);
          y[i + 2] = fmaf((__uint_as_float(r[i + 2]) - mean) * rstd, gm.z, bt.z);
          y[i + 3] = fmaf((__uint_as_float(r[i + 3]) - mean) * rstd, gm.w, bt.w);
        }
        if (with_ffn) {
#pragma unroll
          for (int h8 = 0; h8 < 4; ++h8) {
            const uint4 pk = make_uint4(pack_h2(y[h8 * 8 + 0], y[h8 * 8 + 1]), pack_h2(y[h8 * 8 + 2], y[h8 * 8 + 3]),
                                        pack_h2(y[h8 * 8 + 4], y[h8 * 8 + 5]), pack_h2(y[h8 * 8 + 6], y[h8 * 8 + 7]));
            *reinterpret_cast<uint4*>(sX + (2 + (cq >> 1)) * kTileBytes + tc::sw128_offset(row, (cq & 1) * 32 + h8 * 8)) = pk;
          }
        } else if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(src_row + i));
            *reinterpret_cast<float4*>(out_row + i) = make_float4(s4.x + y[i], s4.y + y[i + 1], s4.z + y[i + 2], s4.w + y[i + 3]);
          }
        }
        tc::tc_fence_before_sync();
        if (with_ffn) {
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&sm.x_ready);
        }
        tc::mbar_arrive(&sm.acc1_empty[b]);
      }
      if (!with_ffn) continue;
      // ---- hidden chunks: H[hb] = fp16(GELU(acc1[b])), this warp's 32 columns
      for (int j = 0; j < kHidChunks; ++j) {
        const uint32_t b = G & 1u, hb = C & 1u;
        tc::mbar_wait(&sm.acc1_full[b], (G >> 1) & 1u);
        tc::tc_fence_after_sync();
        ++G;
        uint32_t r0[16], r1[16];
        const uint32_t acc = lane_addr + kColAcc1 + b * 128 + cq * 32;
        tc::tmem_ld16(acc, r0);
        tc::tmem_ld16(acc + 16, r1);
        tc::mbar_wait(&sm.h_empty[hb], ((C >> 1) & 1u) ^ 1u);      // FFN2 of chunk j - 2 has read H[hb] (long done: no stall in steady state)
        ++C;
        uint32_t pk[16];
        tc::tmem_wait_ld(r0);
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_h2(gelu_erf(__uint_as_float(r0[2 * i])), gelu_erf(__uint_as_float(r0[2 * i + 1])));
        tc::tmem_wait_ld(r1);
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[8 + i] = pack_h2(gelu_erf(__uint_as_float(r1[2 * i])), gelu_erf(__uint_as_float(r1[2 * i + 1])));
        tc::tc_fence_after_sync();
        tc::tmem_st16(lane_addr + kColH + hb * 64 + cq * 16, pk);
        tc::tmem_wait_st();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&sm.h_ready[hb]);
        tc::mbar_arrive(&sm.acc1_empty[b]);
      }
      // ---- LayerNorm 2 on the FFN output + residual
      tc::mbar_wait(&sm.acc2_full, it & 1u);
      tc::tc_fence_after_sync();
      {
        const uint32_t acc = lane_addr + kColAcc2;
        float mean, rstd;
        row_stats(acc, eps, mean, rstd);
        uint32_t r[32];
        tc::tmem_ld32(acc + cq * 32, r);
        tc::tmem_wait_ld(r);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int col = cq * 32 + i;
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(src_row + i));
            const float4 gm = __ldg(reinterpret_cast<const float4*>(ln + 256 + col));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(ln + 384 + col));
            float4 o;
            o.x = s4.x + fmaf((__uint_as_float(r[i + 0]) - mean) * rstd, gm.x, bt.x);
            o.y = s4.y + fmaf((__uint_as_float(r[i + 1]) - mean) * rstd, gm.y, bt.y);
            o.z = s4.z + fmaf((__uint_as_float(r[i + 2]) - mean) * rstd, gm.z, bt.z);
            o.w = s4.w + fmaf((__uint_as_float(r[i + 3]) - mean) * rstd, gm.w, bt.w);
            *reinterpret_cast<float4*>(out_row + i) = o;
          }
        }
        tc::tc_fence_before_sync();      // orders these TMEM reads before the a_ready arrival of the next tile (acc2 is rewritten after it)
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// weights -> fp16 pre-swizzled tiles in streaming order + the LayerNorm parameters; one thread per 16-byte chunk
__global__ void token_block_pack_kernel(const float* __restrict__ merge_w, const float* __restrict__ w1, const float* __restrict__ w2,
                                        const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
                                        const float* __restrict__ b2, unsigned char* __restrict__ outp, const int n_tiles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles * 1024) {
    const int tile = i >> 10, r = (i >> 3) & 127, c8 = i & 7;
    const float* src;
    if (tile < 2) {
      src = merge_w + (size_t)r * kCh + tile * 64 + c8 * 8;
    } else if (tile < 34) {
      const int j = (tile - 2) >> 2, kb = (tile - 2) & 3;
      src = w1 + (size_t)(j * 128 + r) * (2 * kCh) + kb * 64 + c8 * 8;
    } else {
      const int j = (tile - 34) >> 1, kb = (tile - 34) & 1;
      src = w2 + (size_t)r * kHid + j * 128 + kb * 64 + c8 * 8;
    }
    const float4 lo = __ldg(reinterpret_cast<const float4*>(src)), hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
    *reinterpret_cast<uint4*>(outp + (size_t)tile * kTileBytes + tc::sw128_offset(r, c8 * 8)) =
        make_uint4(pack_h2(lo.x, lo.y), pack_h2(lo.z, lo.w), pack_h2(hi.x, hi.y), pack_h2(hi.z, hi.w));
  }
  if (i < 512) {
    float* ln = reinterpret_cast<float*>(outp + (size_t)n_tiles * kTileBytes);
    const int k = i & 127, which = i >> 7;
    const float* p = which == 0 ? g1 : which == 1 ? b1 : which == 2 ? g2 : b2;
    ln[i] = p ? p[k] : (which == 2 ? 1.f : 0.f);
  }
}

PerDevice<int> g_block_sm_count;

}  // namespace

int64_t token_block_weight_bytes(int with_ffn) { return (int64_t)(with_ffn ? kTilesFfn : kTilesNoFfn) * kTileBytes + 512 * 4; }

int launch_token_block_pack(const float* merge_w, const float* g1, const float* b1, const float* w1, const float* w2, const float* g2,
                            const float* b2, void* out, int with_ffn, cudaStream_t s) {
  const int n_tiles = with_ffn ? kTilesFfn : kTilesNoFfn;
  const int threads = n_tiles * 1024;
  token_block_pack_kernel<<<(threads + 255) / 256, 256, 0, s>>>(merge_w, w1, w2, g1, b1, g2, b2, reinterpret_cast<unsigned char*>(out), n_tiles);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

int launch_token_block(const float* attn, const float* source, const void* weights, int with_ffn, float eps, float* out, int64_t T,
                       cudaStream_t s) {
  if (T <= 0) return MNF_OK;
  int& n_sm = g_block_sm_count.cur();
  if (n_sm == 0) {
    int dev = 0;
    MNF_CUDA_TRY(cudaGetDevice(&dev));
    MNF_CUDA_TRY(cudaFuncSetAttribute(token_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    MNF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t n_tiles = (T + kTok - 1) / kTok;
  const unsigned grid = (unsigned)(n_tiles < n_sm ? n_tiles : n_sm);
  token_block_kernel<<<grid, kThreads, kSmemBytes, s>>>(attn, source, reinterpret_cast<const unsigned char*>(weights), out, T, with_ffn, eps);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
