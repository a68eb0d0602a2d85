// K-mlp-composite, tcgen05 variant: conditional MLP on the 5th-gen tensor cores + ray transformer + compositing.
//
// Replaces CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100), MultiHeadAttention.forward
// (models/rfdecoder/ray_transformer.py:49-79), NeRF.composite (models/rfdecoder/nerf.py:101-124) and the view-0
// NDC / ray-direction preparation of MatchNeRF.render (models/matchnerf.py:120-134).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      weight streamer: the packed fp16 weight chunks (pre-swizzled SWIZZLE_128B K-major tiles) are pulled
//               through a 6-stage shared-memory ring with cp.async.bulk (TMA unit) + mbarrier complete_tx;
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::f16 with the ACTIVATIONS AS THE A OPERAND IN TENSOR
//               MEMORY (.ts form) and the weights as the B operand from the ring; accumulators in tensor memory;
//   warps 2-9   two "slots" of 128 threads; a slot owns one 128-sample tile (thread = sample row = TMEM lane).
//               A slot thread stages its sample (geometry -> positional encoding -> fp16 A operand via tcgen05.st),
//               runs every epilogue h = relu((acc + b) * gate) -> fp16 -> tcgen05.st, the 16-wide ray transformer
//               over the samples of its ray, and the alpha compositing scan.  Ray state never leaves the SM.
// The two slots run the same layer in lock step so one streamed weight chunk feeds 256 samples; while slot A's
// accumulator is in its epilogue the tensor pipe works on slot B.
//
// Tensor-memory map per slot (256 columns): D acc [0,128) | H fp16 act [128,192) | ENC fp16 [192,224) | COND fp16 [224,240)
//
// Precision: fp16 operands, fp32 accumulation; gate and the pre-gate sum are rounded to fp16 before their product
// (DESIGN.md "precision").  Activations are assumed to stay below the fp16 range (65504).
#include <cuda_fp16.h>

#include <vector>

#include "decoder_weights.cuh"
#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {

constexpr int kTileM = 128;
constexpr int kNumStages = 6;
constexpr int kChunkBytes = 128 * 128;            // [128 rows][64 fp16]
constexpr int kHeadN = 80;                        // 16 alpha + 64 colour-hidden outputs
constexpr int kHeadChunkBytes = kHeadN * 128;
constexpr int kNumChunks = 15;                    // gate, L0, 4 x 2, 3 (L5), 2 (heads)
constexpr int kNumPhases = 8;                     // gate, L0..L5, heads
constexpr int kEpiThreads = 256;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kColD = 0, kColH = 128, kColEnc = 192, kColCond = 224, kSlotCols = 256;
constexpr int kMaxRaysPerTile = 8;                // S >= 16

__host__ __device__ constexpr int chunk_bytes(int c) { return c < 13 ? kChunkBytes : kHeadChunkBytes; }
__host__ __device__ constexpr int chunk_offset(int c) { return c <= 13 ? c * kChunkBytes : 13 * kChunkBytes + (c - 13) * kHeadChunkBytes; }
constexpr int kPackedBytes = 13 * kChunkBytes + 2 * kHeadChunkBytes;

struct TcParams {             // small fp32 parameters, staged into shared memory once per CTA
  float bias[7][128];         // trunk layers 0..5, [6] = pts_bias.bias
  float att_q[256], att_k[256], att_v[256], att_fc[256];
  float ln_w[16], ln_b[16];
  float oa0_w[256], oa0_b[16], oa2_w[16];
  float alpha_b[16];
  float views_dir[64 * 3];    // views_linears.0.weight[:, 128:131]
  float views_b[64];          // views bias + views_w[:, :128] . feature_linear.bias   (feature_linear folded in)
  float rgb_w[3 * 64];
  float rgb_b[3];
  float oa2_b;
};

struct TcSmem {
  alignas(1024) unsigned char ring[kNumStages][kChunkBytes];
  TcParams p;
  float kbuf[2][kTileM][16];
  float vbuf[2][kTileM][16];
  float dirvec[2][kMaxRaysPerTile][64];
  float red[2][4][8];
  alignas(8) uint64_t w_full[kNumStages];
  uint64_t w_empty[kNumStages];
  uint64_t a_ready[2];
  uint64_t d_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void slot_barrier(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t gate_relu(uint32_t t, uint32_t g) {  // relu(t * g) on packed halves
  const __half2 r = __hfma2_relu(*reinterpret_cast<const __half2*>(&t), *reinterpret_cast<const __half2*>(&g), __float2half2_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ float act_fn(int kind, float x) { return kind == 0 ? fmaxf(x, 0.f) : (x > 0.f ? x : expm1f(x)); }

}  // namespace

struct DecoderWeightsTC {
  unsigned char* packed = nullptr;   // device: kPackedBytes of pre-swizzled fp16 chunks
  TcParams* params = nullptr;        // device
};

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
decoder_tc_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const mnf_decoder_cfg cfg,
                  const unsigned char* __restrict__ wpacked, const TcParams* __restrict__ gparams,
                  const __half* __restrict__ cond, const int setbg_opaque, float* __restrict__ out_rgb,
                  float* __restrict__ out_depth, float* __restrict__ out_opacity, float* __restrict__ aux) {
  extern __shared__ unsigned char smem_dyn[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = cfg.n_samples;
  const int rays_per_tile = kTileM / S;
  const int64_t n_tiles = (rays.n_rays + rays_per_tile - 1) / rays_per_tile;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  // ---- one-time setup
  for (int i = tid; i < (int)(sizeof(TcParams) / 4); i += blockDim.x)
    reinterpret_cast<float*>(&sm.p)[i] = reinterpret_cast<const float*>(gparams)[i];
  if (tid == 0) {
    for (int i = 0; i < kNumStages; ++i) {
      tc::mbar_init(&sm.w_full[i], 1);
      tc::mbar_init(&sm.w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sm.a_ready[i], kTileM);
      tc::mbar_init(&sm.d_full[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ================================================================== weight streamer
    if (lane == 0) {
      uint32_t n = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int c = 0; c < kNumChunks; ++c, ++n) {
          const uint32_t st = n % kNumStages, par = (n / kNumStages) & 1;
          tc::mbar_wait(&sm.w_empty[st], par ^ 1);
          tc::mbar_arrive_expect_tx(&sm.w_full[st], chunk_bytes(c));
          tc::bulk_g2s(sm.ring[st], wpacked + chunk_offset(c), chunk_bytes(c), &sm.w_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc128 = tc::umma_idesc_f16(128, 128), idesc_head = tc::umma_idesc_f16(128, kHeadN);
      uint32_t n = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const bool active1 = 2 * pair + 1 < n_tiles;
#pragma unroll 1
        for (int ph = 0; ph < kNumPhases; ++ph) {
          const int nch = ph <= 1 ? 1 : (ph == 6 ? 3 : 2);
          for (int j = 0; j < nch; ++j) tc::mbar_wait(&sm.w_full[(n + j) % kNumStages], ((n + j) / kNumStages) & 1);
#pragma unroll 1
          for (int slot = 0; slot < 2; ++slot) {
            if (slot == 1 && !active1) break;
            tc::mbar_wait(&sm.a_ready[slot], ph & 1);
            tc::tc_fence_after_sync();
            const uint32_t tb = tmem + slot * kSlotCols;
            const uint32_t d = tb + kColD;
            auto bdesc = [&](int j, int ks) { return tc::umma_desc_sw128(tc::smem_u32(sm.ring[(n + j) % kNumStages]) + ks * 32); };
            if (ph == 0) {          // gate = pts_bias(cond): K = 32
              for (int ks = 0; ks < 2; ++ks) tc::umma_ts(d, tb + kColCond + ks * 8, bdesc(0, ks), idesc128, ks > 0);
            } else if (ph == 1) {   // layer 0: K = 64 (encoding)
              for (int ks = 0; ks < 4; ++ks) tc::umma_ts(d, tb + kColEnc + ks * 8, bdesc(0, ks), idesc128, ks > 0);
            } else if (ph <= 5) {   // layers 1..4: K = 128
              for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(ks >> 2, ks & 3), idesc128, ks > 0);
            } else if (ph == 6) {   // layer 5 on [enc, h]: K = 64 + 128
              for (int ks = 0; ks < 4; ++ks) tc::umma_ts(d, tb + kColEnc + ks * 8, bdesc(0, ks), idesc128, ks > 0);
              for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(1 + (ks >> 2), ks & 3), idesc128, 1);
            } else {                // heads: [alpha_linear | views(feature_linear(.))], N = 80, K = 128
              for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(ks >> 2, ks & 3), idesc_head, ks > 0);
            }
            tc::umma_commit(&sm.d_full[slot]);
          }
          for (int j = 0; j < nch; ++j) tc::umma_commit(&sm.w_empty[(n + j) % kNumStages]);
          n += nch;
        }
      }
    }
  } else {
    // ================================================================== slot threads (epilogue / ray state)
    const int slot = (warp - 2) >> 2;
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;             // sample row inside the tile == TMEM lane
    const uint32_t tb = tmem + slot * kSlotCols + ((uint32_t)(quarter * 32) << 16);
    const int ray_local = row / S, s = row - ray_local * S;
    const float kLog2e = 1.4426950408889634f;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile = 2 * pair + slot;
      if (tile >= n_tiles) break;
      const int64_t ray = tile * rays_per_tile + ray_local;
      const bool valid = ray < rays.n_rays;
      const size_t n_glob = valid ? (size_t)ray * S + s : 0;

      // ---------------- stage this sample: geometry -> positional encoding -> fp16 A operands in tensor memory
      float depth_t = 0.f, n_views_seen = 0.f;
      {
        float x[3] = {0.f, 0.f, 0.f};
        float dir[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t pix = rays.ray_idx ? rays.ray_idx[ray] : rays.first_ray + ray;
          float o[3], d[3];
          cast_ray(cams, pix, o, d);
          const float u = rays.jitter ? rays.jitter[n_glob] : 0.f;
          depth_t = sample_depth(cams, s, S, u);
          float p[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], depth_t));
          project_ndc(cams, 0, p, x[0], x[1], x[2]);
          const float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
          const float ux = d[0] / nrm, uy = d[1] / nrm, uz = d[2] / nrm;
          const float* E = cams.w2c[0];
          dir[0] = ux * E[0] + uy * E[1] + uz * E[2];
          dir[1] = ux * E[4] + uy * E[5] + uz * E[6];
          dir[2] = ux * E[8] + uy * E[9] + uz * E[10];
        }
        // direction term of the colour head, one vector per ray (the ray's threads split its 64 outputs)
        for (int o2 = s; o2 < 64; o2 += S)
          sm.dirvec[slot][ray_local][o2] = sm.p.views_dir[o2 * 3] * dir[0] + sm.p.views_dir[o2 * 3 + 1] * dir[1] +
                                           sm.p.views_dir[o2 * 3 + 2] * dir[2] + sm.p.views_b[o2];
        // encoding order (cond_nerf.py:108-116, :56-57): x, sin(2^k x) k-major, cos(2^k x) k-major, zero pad
        float sn[3], cs[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) sincosf(x[i], &sn[i], &cs[i]);
        uint32_t e[32];
        float prev = 0.f;  // pairs are emitted in index order: idx 0..63
        auto put = [&](int idx, float v) {
          if (idx & 1) e[idx >> 1] = pack_h2(prev, v); else prev = v;
        };
        put(0, x[0]); put(1, x[1]); put(2, x[2]);
        float sk[3] = {sn[0], sn[1], sn[2]}, ck[3] = {cs[0], cs[1], cs[2]};
        float cosv[30];
#pragma unroll
        for (int k = 0; k < kL3D; ++k) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            put(3 + k * 3 + i, sk[i]);
            cosv[k * 3 + i] = ck[i];
            const float s2 = 2.f * sk[i] * ck[i];          // double-angle recurrence: 2^k x -> 2^(k+1) x
            const float c2 = 1.f - 2.f * sk[i] * sk[i];
            sk[i] = s2; ck[i] = c2;
          }
        }
#pragma unroll
        for (int j = 0; j < 30; ++j) put(33 + j, cosv[j]);
        put(63, 0.f);
        {
          uint32_t lo[16], hi[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { lo[j] = e[j]; hi[j] = e[16 + j]; }
          tc::tmem_st16(tb + kColEnc, lo);
          tc::tmem_st16(tb + kColEnc + 16, hi);
        }
        uint32_t cnd[16];
        if (valid) {
          const uint4* src = reinterpret_cast<const uint4*>(cond + n_glob * kCondPad);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 v4 = __ldg(src + j);
            cnd[4 * j] = v4.x; cnd[4 * j + 1] = v4.y; cnd[4 * j + 2] = v4.z; cnd[4 * j + 3] = v4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) cnd[j] = 0u;
        }
        {  // visibility masks live at cond[19..21]
          const __half2 h9 = *reinterpret_cast<const __half2*>(&cnd[9]);
          const __half2 h10 = *reinterpret_cast<const __half2*>(&cnd[10]);
          n_views_seen = __high2float(h9) + __low2float(h10) + __high2float(h10);
        }
        tc::tmem_st16(tb + kColCond, cnd);
        tc::tmem_wait_st();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&sm.a_ready[slot]);
      }

      // ---------------- gate = pts_bias(cond) + b, kept as 64 packed-half registers for all six layers
      uint32_t gate[64];
      tc::mbar_wait(&sm.d_full[slot], 0);
      tc::tc_fence_after_sync();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tb + kColD + c0, r);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          gate[c0 / 2 + j] = pack_h2(__uint_as_float(r[2 * j]) + sm.p.bias[6][c0 + 2 * j], __uint_as_float(r[2 * j + 1]) + sm.p.bias[6][c0 + 2 * j + 1]);
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(&sm.a_ready[slot]);

      // ---------------- trunk: h = relu((acc + b_l) * gate) -> fp16 -> tensor memory (next layer's A operand)
#pragma unroll 1
      for (int l = 0; l < kDepth; ++l) {
        tc::mbar_wait(&sm.d_full[slot], (l + 1) & 1);
        tc::tc_fence_after_sync();
        const float* bl = sm.p.bias[l];
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tc::tmem_ld32(tb + kColD + c0, r);
          tc::tmem_wait_ld();
          uint32_t o16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 b2 = *reinterpret_cast<const float2*>(bl + c0 + 2 * j);
            o16[j] = gate_relu(pack_h2(__uint_as_float(r[2 * j]) + b2.x, __uint_as_float(r[2 * j + 1]) + b2.y), gate[c0 / 2 + j]);
          }
          tc::tmem_st16(tb + kColH + c0 / 2, o16);
        }
        tc::tmem_wait_st();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&sm.a_ready[slot]);
      }

      // ---------------- heads: raw alpha (16) and the colour hidden layer (64)
      float xr[16];
      float rgb[3];
      tc::mbar_wait(&sm.d_full[slot], 1);
      tc::tc_fence_after_sync();
      {
        uint32_t r16[16];
        tc::tmem_ld16(tb + kColD, r16);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v = act_fn(cfg.raytrans_act, __uint_as_float(r16[j]) + sm.p.alpha_b[j]);
          if (cfg.raytrans_posenc) {   // cond_nerf.py:118-127
            const float ang = (float)s * exp2f(-(float)(j >> 1) * (13.287712379549449f / 8.f));   // s / 10000^(2*(j/2)/16)
            v += (j & 1) ? cosf(ang) : sinf(ang);
          }
          xr[j] = v;
        }
        float acc3[3] = {sm.p.rgb_b[0], sm.p.rgb_b[1], sm.p.rgb_b[2]};
#pragma unroll
        for (int c0 = 16; c0 < 80; c0 += 32) {
          uint32_t r[32];
          tc::tmem_ld32(tb + kColD + c0, r);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int o2 = c0 - 16 + j;
            const float hv = fmaxf(__uint_as_float(r[j]) + sm.dirvec[slot][ray_local][o2], 0.f);
            acc3[0] = fmaf(hv, sm.p.rgb_w[o2], acc3[0]);
            acc3[1] = fmaf(hv, sm.p.rgb_w[64 + o2], acc3[1]);
            acc3[2] = fmaf(hv, sm.p.rgb_w[128 + o2], acc3[2]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = 1.f / (1.f + __expf(-acc3[c]));
      }

      // ---------------- ray transformer over the S samples of this ray (ray_transformer.py:49-79)
      float q[16];
#pragma unroll
      for (int oi = 0; oi < 16; ++oi) {
        float aq = 0.f, ak = 0.f, av = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          aq = fmaf(xr[i], sm.p.att_q[oi * 16 + i], aq);
          ak = fmaf(xr[i], sm.p.att_k[oi * 16 + i], ak);
          av = fmaf(xr[i], sm.p.att_v[oi * 16 + i], av);
        }
        q[oi] = aq * (0.5f * kLog2e);          // 1/temperature (sqrt(d_k) = 2), and log2(e) for exp2
        sm.kbuf[slot][row][oi] = ak;
        sm.vbuf[slot][row][oi] = av;
      }
      slot_barrier(slot);
      float sigma;
      {
        const bool row_valid = n_views_seen > 1.f;    // cond_nerf.py:83; masks whole QUERY rows (uniform attention)
        const float4* kb = reinterpret_cast<const float4*>(&sm.kbuf[slot][ray_local * S][0]);
        const float4* vb = reinterpret_cast<const float4*>(&sm.vbuf[slot][ray_local * S][0]);
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (row_valid) {
          for (int j = 0; j < S; ++j) {
#pragma unroll
            for (int hd = 0; hd < 4; ++hd) {
              const float4 k4 = kb[j * 4 + hd];
              const float sc = fmaf(q[hd * 4 + 3], k4.w, fmaf(q[hd * 4 + 2], k4.z, fmaf(q[hd * 4 + 1], k4.y, q[hd * 4] * k4.x)));
              mx[hd] = fmaxf(mx[hd], sc);
            }
          }
        } else {
#pragma unroll
          for (int hd = 0; hd < 4; ++hd) { mx[hd] = 0.f; }
#pragma unroll
          for (int i = 0; i < 16; ++i) q[i] = 0.f;
        }
        float den[4] = {0.f, 0.f, 0.f, 0.f}, att[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) att[i] = 0.f;
        for (int j = 0; j < S; ++j) {
#pragma unroll
          for (int hd = 0; hd < 4; ++hd) {
            const float4 k4 = kb[j * 4 + hd];
            const float4 v4 = vb[j * 4 + hd];
            const float sc = fmaf(q[hd * 4 + 3], k4.w, fmaf(q[hd * 4 + 2], k4.z, fmaf(q[hd * 4 + 1], k4.y, fmaf(q[hd * 4], k4.x, -mx[hd]))));
            const float pe = exp2f(sc);
            den[hd] += pe;
            att[hd * 4 + 0] = fmaf(pe, v4.x, att[hd * 4 + 0]);
            att[hd * 4 + 1] = fmaf(pe, v4.y, att[hd * 4 + 1]);
            att[hd * 4 + 2] = fmaf(pe, v4.z, att[hd * 4 + 2]);
            att[hd * 4 + 3] = fmaf(pe, v4.w, att[hd * 4 + 3]);
          }
        }
#pragma unroll
        for (int hd = 0; hd < 4; ++hd) {
          const float inv = 1.f / den[hd];
#pragma unroll
          for (int dd = 0; dd < 4; ++dd) att[hd * 4 + dd] *= inv;
        }
        float y[16], mu = 0.f;
#pragma unroll
        for (int oi = 0; oi < 16; ++oi) {
          float a = xr[oi];
#pragma unroll
          for (int i = 0; i < 16; ++i) a = fmaf(att[i], sm.p.att_fc[oi * 16 + i], a);
          y[oi] = a;
          mu += a;
        }
        mu *= (1.f / 16.f);
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) var += (y[i] - mu) * (y[i] - mu);
        const float rstd = rsqrtf(var * (1.f / 16.f) + 1e-6f);
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = (y[i] - mu) * rstd * sm.p.ln_w[i] + sm.p.ln_b[i];
        float acc = sm.p.oa2_b;
#pragma unroll
        for (int oi = 0; oi < 16; ++oi) {
          float a = sm.p.oa0_b[oi];
#pragma unroll
          for (int i = 0; i < 16; ++i) a = fmaf(y[i], sm.p.oa0_w[oi * 16 + i], a);
          acc = fmaf(act_fn(cfg.raytrans_act, a), sm.p.oa2_w[oi], acc);
        }
        sigma = fmaxf(acc, 0.f);
        if (cfg.density_maskfill && n_views_seen < 1.f) sigma = 0.f;
        if (!valid) sigma = 0.f;
      }
      if (aux && valid) {
        float4* a4 = reinterpret_cast<float4*>(aux) + n_glob;
        *a4 = make_float4(rgb[0], rgb[1], rgb[2], sigma);
      }

      // ---------------- alpha compositing along the ray (nerf.py:101-124, wo_render_interval)
      {
        const int seg = S < 32 ? S : 32;            // rows of one ray inside a warp
        float incl = sigma;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float nb = __shfl_up_sync(0xffffffffu, incl, off, 32);
          if ((lane & (seg - 1)) >= off && off < seg) incl += nb;
        }
        const int wq = quarter;                     // warp index inside the slot == row / 32
        if (S > 32) {
          if (lane == 31) sm.red[slot][wq][5] = incl;
          slot_barrier(slot);
          const int w0 = (ray_local * S) >> 5;      // first warp of this ray
          float base = 0.f;
          for (int w2 = w0; w2 < wq; ++w2) base += sm.red[slot][w2][5];
          incl += base;
        }
        const float excl = incl - sigma;
        const float wgt = valid ? __expf(-excl) * (1.f - __expf(-sigma)) : 0.f;
        float part[5] = {wgt * rgb[0], wgt * rgb[1], wgt * rgb[2], wgt * depth_t, wgt};
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1)
            if (off < seg) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
        if (S > 32) {
          if (lane == 0)
#pragma unroll
            for (int i = 0; i < 5; ++i) sm.red[slot][wq][i] = part[i];
          slot_barrier(slot);
          if (s == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              float t = 0.f;
              for (int w2 = 0; w2 < S / 32; ++w2) t += sm.red[slot][wq + w2][i];
              part[i] = t;
            }
          }
        }
        if (s == 0 && valid) {
          const float bg = setbg_opaque ? 1.f - part[4] : 0.f;
          out_rgb[ray * 3 + 0] = part[0] + bg;
          out_rgb[ray * 3 + 1] = part[1] + bg;
          out_rgb[ray * 3 + 2] = part[2] + bg;
          out_depth[ray] = part[3];
          out_opacity[ray] = part[4];
        }
        slot_barrier(slot);   // kbuf / vbuf / red / dirvec are rewritten by the next tile
      }
    }
  }

  // ---- teardown
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// host side: weight packing and launch
// ------------------------------------------------------------------------------------------------------------------
namespace {

void put_chunk(std::vector<__half>& buf, int chunk, int n_rows, const std::vector<double>& W /*[n_rows][64]*/) {
  __half* base = buf.data() + chunk_offset(chunk) / 2;
  for (int n = 0; n < n_rows; ++n)
    for (int k = 0; k < 64; ++k) base[tc::sw128_offset(n, k) / 2] = __float2half_rn((float)W[(size_t)n * 64 + k]);
}

}  // namespace

int decoder_tc_pack(const float* P, const ParamOffsets& off, DecoderWeightsTC** out) {
  std::vector<__half> buf(kPackedBytes / 2, __float2half_rn(0.f));
  auto slice = [&](const float* W, int N, int K, int k0, int kn) {   // W[N][K] columns [k0, k0+kn) -> [N][64]
    std::vector<double> t((size_t)N * 64, 0.0);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < kn; ++k) t[(size_t)n * 64 + k] = W[(size_t)n * K + k0 + k];
    return t;
  };
  put_chunk(buf, 0, 128, slice(P + off.gate_w, 128, kCond, 0, kCond));
  put_chunk(buf, 1, 128, slice(P + off.pts_w[0], 128, kEnc, 0, kEnc));
  for (int l = 1; l <= 4; ++l) {
    put_chunk(buf, 2 * l, 128, slice(P + off.pts_w[l], 128, 128, 0, 64));
    put_chunk(buf, 2 * l + 1, 128, slice(P + off.pts_w[l], 128, 128, 64, 64));
  }
  put_chunk(buf, 10, 128, slice(P + off.pts_w[5], 128, kEnc + 128, 0, kEnc));
  put_chunk(buf, 11, 128, slice(P + off.pts_w[5], 128, kEnc + 128, kEnc, 64));
  put_chunk(buf, 12, 128, slice(P + off.pts_w[5], 128, kEnc + 128, kEnc + 64, 64));
  // heads: rows 0..15 alpha_linear; rows 16..79 views_linears.0[:, :128] @ feature_linear (no nonlinearity between them)
  std::vector<double> head((size_t)kHeadN * 128, 0.0);
  for (int n = 0; n < 16; ++n)
    for (int k = 0; k < 128; ++k) head[(size_t)n * 128 + k] = P[off.alpha_w + (size_t)n * 128 + k];
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  for (int n = 0; n < 64; ++n) {
    const float* wv = P + off.views_w + (size_t)n * (128 + 3);
    for (int k = 0; k < 128; ++k) {
      double a = 0.0;
      for (int j = 0; j < 128; ++j) a += (double)wv[j] * (double)P[off.feat_w + (size_t)j * 128 + k];
      head[(size_t)(16 + n) * 128 + k] = a;
    }
    double b = P[off.views_b + n];
    for (int j = 0; j < 128; ++j) b += (double)wv[j] * (double)P[off.feat_b + j];
    tp.views_b[n] = (float)b;
    for (int k = 0; k < 3; ++k) tp.views_dir[n * 3 + k] = wv[128 + k];
  }
  for (int half = 0; half < 2; ++half) {
    std::vector<double> t((size_t)kHeadN * 64);
    for (int n = 0; n < kHeadN; ++n)
      for (int k = 0; k < 64; ++k) t[(size_t)n * 64 + k] = head[(size_t)n * 128 + half * 64 + k];
    put_chunk(buf, 13 + half, kHeadN, t);
  }
  for (int l = 0; l < kDepth; ++l) memcpy(tp.bias[l], P + off.pts_b[l], 128 * sizeof(float));
  memcpy(tp.bias[6], P + off.gate_b, 128 * sizeof(float));
  memcpy(tp.att_q, P + off.att_q, sizeof(tp.att_q));
  memcpy(tp.att_k, P + off.att_k, sizeof(tp.att_k));
  memcpy(tp.att_v, P + off.att_v, sizeof(tp.att_v));
  memcpy(tp.att_fc, P + off.att_fc, sizeof(tp.att_fc));
  memcpy(tp.ln_w, P + off.ln_w, sizeof(tp.ln_w));
  memcpy(tp.ln_b, P + off.ln_b, sizeof(tp.ln_b));
  memcpy(tp.oa0_w, P + off.oa0_w, sizeof(tp.oa0_w));
  memcpy(tp.oa0_b, P + off.oa0_b, sizeof(tp.oa0_b));
  memcpy(tp.oa2_w, P + off.oa2_w, sizeof(tp.oa2_w));
  tp.oa2_b = P[off.oa2_b];
  memcpy(tp.alpha_b, P + off.alpha_b, sizeof(tp.alpha_b));
  memcpy(tp.rgb_w, P + off.rgb_w, sizeof(tp.rgb_w));
  memcpy(tp.rgb_b, P + off.rgb_b, sizeof(tp.rgb_b));

  DecoderWeightsTC* w = new DecoderWeightsTC();
  MNF_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&w->packed), kPackedBytes));
  MNF_CUDA_TRY(cudaMemcpy(w->packed, buf.data(), kPackedBytes, cudaMemcpyHostToDevice));
  MNF_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&w->params), sizeof(TcParams)));
  MNF_CUDA_TRY(cudaMemcpy(w->params, &tp, sizeof(TcParams), cudaMemcpyHostToDevice));
  *out = w;
  return MNF_OK;
}

void decoder_tc_free(DecoderWeightsTC* w) {
  if (!w) return;
  cudaFree(w->packed);
  cudaFree(w->params);
  delete w;
}

bool decoder_tc_supports(const mnf_decoder_cfg& cfg) {
  const int S = cfg.n_samples;
  return S == 16 || S == 32 || S == 64 || S == 128;
}

int launch_decoder_tc(const DevCams& cams, const DevRays& rays, const mnf_decoder_cfg& cfg, const DecoderWeightsTC* w,
                      const HeadParams*, const __half* cond_f16, int setbg_opaque, float* out_rgb, float* out_depth,
                      float* out_opacity, float* aux, cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  if (!w) { set_error("tcgen05 decoder weights not packed"); return MNF_ESTATE; }
  static int n_sm = 0;
  const size_t smem = sizeof(TcSmem) + 1024;
  if (n_sm == 0) {
    int dev = 0;
    MNF_CUDA_TRY(cudaGetDevice(&dev));
    MNF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    MNF_CUDA_TRY(cudaFuncSetAttribute(decoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int rays_per_tile = kTileM / cfg.n_samples;
  const int64_t n_tiles = (rays.n_rays + rays_per_tile - 1) / rays_per_tile;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const unsigned grid = (unsigned)(n_pairs < n_sm ? n_pairs : n_sm);
  decoder_tc_kernel<<<grid, kThreads, smem, s>>>(cams, rays, cfg, w->packed, w->params, cond_f16, setbg_opaque, out_rgb,
                                                 out_depth, out_opacity, aux);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
