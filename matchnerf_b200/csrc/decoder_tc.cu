// K-mlp-composite, tcgen05 variant: conditional MLP on the 5th-gen tensor cores + ray transformer + compositing.
//
// Replaces CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100), MultiHeadAttention.forward
// (models/rfdecoder/ray_transformer.py:49-79), NeRF.composite (models/rfdecoder/nerf.py:101-124) and the view-0
// NDC / ray-direction preparation of MatchNeRF.render (models/matchnerf.py:120-134).
//
// Structure (one persistent CTA per SM, 640 threads = 5 warpgroups, registers rebalanced with setmaxnreg):
//   warp 0      weight streamer: the packed fp16 weight chunks (pre-swizzled SWIZZLE_128B K-major tiles) are pulled
//               through a 6-stage shared-memory ring with cp.async.bulk (TMA unit) + mbarrier complete_tx;
//   warp 1      MMA issuer: one thread issues tcgen05.mma.kind::f16 with the ACTIVATIONS AS THE A OPERAND IN TENSOR
//               MEMORY (.ts form) and the weights as the B operand from the ring; accumulators in tensor memory;
//   warps 4-11  two TRUNK slots of 128 threads (136 registers); a slot owns one 128-sample tile (thread = sample row
//               = TMEM lane).  A trunk thread stages its sample (geometry -> positional encoding -> fp16 A operand via
//               tcgen05.st), runs every epilogue h = relu((acc + b) * gate) -> fp16 -> tcgen05.st and the colour /
//               alpha heads, then hands 21 floats per sample to its ray group through shared memory;
//   warps 12-19 two RAY groups of 128 threads (88 registers): the 16-wide 4-head ray transformer over the samples of a
//               ray, the density head and the alpha-compositing scan.  Ray state never leaves the SM.
// The two trunk slots run the same layer in lock step so one streamed weight chunk feeds 256 samples; while slot A's
// accumulator is in its epilogue the tensor pipe works on slot B, and while the trunk warps wait for the tensor pipe
// the ray groups (decoupled by a double-buffered hand-off) keep the CUDA cores busy with the previous tiles.
//
// Tensor-memory map per slot (256 columns): D acc [0,128) | H fp16 act [128,192) | ENC fp16 [192,224) | COND fp16 [224,240)
//
// Precision: fp16 operands, fp32 accumulation; gate and the pre-gate sum are rounded to fp16 before their product
// (DESIGN.md "precision").  Activations are assumed to stay below the fp16 range (65504).
#include <cuda_fp16.h>

#include <cmath>
#include <vector>

#include "decoder_weights.cuh"
#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {

constexpr int kTileM = 128;
constexpr int kNumStages = 4;                      // weight ring depth (16 KB each); 6 -> 4 pays for the staged-row buffers below
constexpr int kChunkBytes = 128 * 128;            // [128 rows][64 fp16]
constexpr int kWRow = 24;                         // halves per row of the small fp16 matrices / of kbuf (48 B)
constexpr int kVRow = 40;                         // halves per key of vbuf (4 heads x 8 columns + pad = 80 B)
constexpr int kHandRow = 20;                      // floats per row of the trunk -> ray hand-off (80 B)
constexpr int kStageRow = 52;                     // 32-bit words per staged sample row: ENC 32 | COND 16 | depth | pad (208 B: conflict-free 16-byte accesses)
constexpr int kHeadN = 80;                        // 16 alpha + 64 colour-hidden outputs
constexpr int kHeadChunkBytes = kHeadN * 128;
constexpr int kNumChunks = 15;                    // gate, L0, 4 x 2, 3 (L5), 2 (heads)
constexpr int kNumPhases = 8;                     // gate, L0..L5, heads
constexpr int kThreads = 640;                     // 5 warpgroups: control | trunk A | trunk B | ray A | ray B
// register budgets after setmaxnreg.  The kernel launches with 640 x 96 = 61440 registers and the rebalancing must fit
// inside that per-CTA pool: 128*32 + 256*136 + 256*88 = 61440.
constexpr int kRegsCtrl = 32, kRegsTrunk = 136, kRegsRay = 88;
constexpr int kColD = 0, kColH = 128, kColEnc = 192, kColCond = 224, kSlotCols = 256;
constexpr int kMaxRaysPerTile = 8;                // S >= 16
#ifndef MNF_STAGE_GEOM_LAYER
#define MNF_STAGE_GEOM_LAYER 1
#endif
#ifndef MNF_STAGE_ENC_LAYER
#define MNF_STAGE_ENC_LAYER 2
#endif
constexpr int kStageGeomLayer = MNF_STAGE_GEOM_LAYER, kStageEncLayer = MNF_STAGE_ENC_LAYER;   // trunk layers after whose epilogue the next tile is staged

__host__ __device__ constexpr int chunk_bytes(int c) { return c < 13 ? kChunkBytes : kHeadChunkBytes; }
__host__ __device__ constexpr int chunk_offset(int c) { return c <= 13 ? c * kChunkBytes : 13 * kChunkBytes + (c - 13) * kHeadChunkBytes; }   // c = 15: the bias tile
constexpr int kBiasOffset = 13 * kChunkBytes + 2 * kHeadChunkBytes;   // resident bias tile (layers 1..4), not part of the ring stream
constexpr int kPackedBytes = kBiasOffset + kChunkBytes;

struct TcParams {             // small fp32 parameters, staged into shared memory once per CTA
  // ray-transformer matrices as mma.sync B operands: fp16, [out][in] rows padded to kWRow halves (48 B: the 8 rows a
  // quad-group reads land in distinct banks).  rows 0..15 = w_qs * (log2(e)/2), 16..31 = w_ks, 32..47 = w_vs
  __half wqkv_h[48][kWRow];
  __half wqkv_l[48][kWRow];   // fp16 remainder W - fp16(W): the q/k/v projection runs as a 3-term split product (~fp32 accurate)
  __half fc_h[16][kWRow];     // ray_attention.fc
  __half oa0_h[16][kWRow];    // out_alpha_linear.0
  float ln_w[16], ln_b[16];
  float oa0_b[16], oa2_w[16];
  float alpha_b[16];
  float views_dir[64 * 3];    // views_linears.0.weight[:, 128:131]
  float views_b[64];          // views bias + views_w[:, :128] . feature_linear.bias   (feature_linear folded in)
  float rgb_w4[64][4];        // [hidden][r, g, b, 0]
  float rgb_b[3];
  float oa2_b;
};

struct TraceCtl { unsigned long long* buf; unsigned per; };
struct TcSmem {
  alignas(1024) unsigned char ring[kNumStages][kChunkBytes];
  // Biases ride on the tensor pipe.  The positional-encoding tile carries 1.0 in its pad column 63 and the conditioning
  // tile in its pad column 22, so gate / layer 0 / layer 5 (which consume those tiles anyway) find their bias as one
  // more weight column.  Layers 1..4 (K = 128, no spare column) get one extra K=16 step: A = encoding columns 48..63,
  // B = columns 16j..16j+15 of this resident [128][64] tile, zero except column 16j+15 = bias of layer 1+j.
  alignas(1024) unsigned char biasw[kChunkBytes];
  TcParams p;
  alignas(16) __half kbuf[2][kTileM][kWRow];                   // keys, [row][head*4 + dim]
  alignas(16) __half vbuf[2][kTileM][kVRow];                   // values, [row][head][8]: even heads (v0..v3, 1, 0, 0, 0), odd heads (1, 0, 0, 0, v0..v3)
  float sig[2][kTileM];
  float dirvec[2][2][kMaxRaysPerTile][64];                     // [slot][tile parity]: direction term of the colour head
  alignas(16) float4 rayrec[2][2][kMaxRaysPerTile];            // [slot][tile parity][ray]: d(q)/d(depth) of the view-0 projection (x, y, z), valid flag
  alignas(16) uint32_t stage[2][kTileM][kStageRow];            // [slot][row]: the NEXT tile's A operands, staged while this tile's MMAs run
  alignas(16) float red[2][2][4][8];                           // [tile parity][slot][warp]: compositing partials of a ray segment (5 sums, optical depth)
  alignas(16) float hand[2][2][kTileM][kHandRow];              // trunk -> ray hand-off: [slot][buffer][row][raw alpha 0..15, r, g, b, depth]
  float hand_nv[2][2][kTileM];                     // views that see the sample
  alignas(8) uint64_t w_full[kNumStages];
  uint64_t w_empty[kNumStages];
  uint64_t a_ready[2];
  uint64_t bias_full;
  uint64_t d_full[2];
  uint64_t ray_full[2][2];
  uint64_t ray_empty[2][2];
  uint64_t geo_full[2][2];                         // [slot][tile parity]: ray records + dirvec of a tile written (geometry warp -> trunk slot)
  uint64_t geo_empty[2][2];                        // ... and consumed (trunk slot -> geometry warp)
  uint32_t tmem_base;
  TraceCtl trace;
};

// ---- optional timeline trace (debug aid; mnf_debug_decoder_trace): CTA 0 records clock64() at protocol points
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ unsigned int g_trace_cap = 0;
// role: 0 mma, 1 trunk, 2 ray.  Lane 0 of every warp records into that warp's own twentieth of the buffer (no atomics,
// so a probe costs a clock read and one store); the slot field carries slot | quarter << 2.
// The armed buffer is latched into shared memory at kernel start: a probe must not cost a global-memory round trip
// (it sits on the MMA issuer's critical path).
[[maybe_unused]] __device__ __forceinline__ void trace(const TraceCtl& tc_, int role, int slot, int ev, unsigned it, unsigned& n) {
  if (tc_.per != 0u) {
    if (n < tc_.per)
      tc_.buf[(threadIdx.x >> 5) * tc_.per + n] = ((unsigned long long)clock64() << 24) | ((unsigned long long)(role & 15) << 20) |
                                                 ((unsigned long long)(slot & 15) << 16) | ((unsigned long long)(ev & 255) << 8) | (it & 255);
    ++n;
  }
}
#ifdef MNF_DECODER_TRACE
#define TRACE_TRUNK(ev) do { if (lane == 0) trace(sm.trace, 1, slot | (quarter << 2), ev, it, trace_n); } while (0)
#define TRACE_RAY(ev) do { if (lane == 0) trace(sm.trace, 2, slot | (quarter << 2), ev, it, trace_n); } while (0)
#define TRACE_MMA(sl, ev) do { if (leader) trace(sm.trace, 0, sl, ev, (unsigned)(n & 255), trace_m); } while (0)
#else   // probes compile to nothing unless the library is built with -DMNF_DECODER_TRACE (tools/decoder_trace.py)
#define TRACE_TRUNK(ev) do { (void)trace_n; } while (0)
#define TRACE_RAY(ev) do { (void)trace_n; } while (0)
#define TRACE_MMA(sl, ev) do { } while (0)
#endif

[[maybe_unused]] __device__ __forceinline__ void trunk_barrier(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory"); }
__device__ __forceinline__ void ray_barrier(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 3) : "memory"); }
// S = 256: a ray spans BOTH slots of a tile pair (slot 0 = samples 0..127, slot 1 = samples 128..255), so the two ray groups
// share keys / values and the compositing scan: their barriers then span both groups (256 threads, barrier id 5)
template <bool kBoth>
__device__ __forceinline__ void ray_barrier_s(int slot) {
  if constexpr (kBoth) asm volatile("bar.sync 5, 256;" ::: "memory");
  else ray_barrier(slot);
}
// S = 64: a ray is exactly two warps of a ray group, so its keys / values / scan partials only need a 64-thread barrier
// (ids 6..9 = slot * 2 + warp pair): the two rays of a tile no longer wait for each other
__device__ __forceinline__ void ray_pair_barrier(int slot, int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(6 + slot * 2 + (quarter >> 1)) : "memory");
}
using tc::mbar_wait_sleep;
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float h2_lo(uint32_t v) { return __low2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ float h2_hi(uint32_t v) { return __high2float(*reinterpret_cast<const __half2*>(&v)); }
__device__ __forceinline__ uint32_t gate_relu(uint32_t t, uint32_t g) {  // relu(t * g) on packed halves
  const __half2 r = __hfma2_relu(*reinterpret_cast<const __half2*>(&t), *reinterpret_cast<const __half2*>(&g), __float2half2_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <int kAct>
__device__ __forceinline__ float act_fn(float x) {
  if constexpr (kAct == 0) return fmaxf(x, 0.f);
  else return x > 0.f ? x : expm1f(x);
}
// Packed pairs of fp32 kept in ONE 64-bit register for their whole life, so that FFMA2 / FADD2 need no re-packing
// moves (ncu of v3: 25 MOVs per 32 FFMA2 in the attention loop when pairs were rebuilt from scalar registers).
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk(float lo, float hi) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float pk_lo(pk2 v) { return __uint_as_float((uint32_t)(v & 0xffffffffull)); }
__device__ __forceinline__ float pk_hi(pk2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ pk2 pk_fma(pk2 a, pk2 b, pk2 c) {
  pk2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ex2_fast(float x) {   // MUFU.EX2 without the denormal-range fix-up of exp2f()
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- warp-level tensor-core pieces of the ray transformer (mma.sync m16n8k16, fp16 operands, fp32 accumulate).
// Fragment coordinates (g = lane >> 2, t = lane & 3):  A: a0 (row g, k 2t..2t+1), a1 (row g+8, same k), a2 (row g, k 8+2t..),
// a3 (row g+8, k 8+2t..);  B: b0 (k 2t..2t+1, n g), b1 (k 8+2t.., n g);  C: c0,c1 (row g, n 2t, 2t+1), c2,c3 (row g+8, same n).
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// m16n8k8: A: a0 (row g, k 2t..2t+1), a1 (row g+8, same k);  B: b0 (k 2t..2t+1, n g);  C as above
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// sin / cos of the base angle of the positional encoding.  Two-term Cody-Waite reduction to [-pi, pi] followed by the MUFU
// approximations (abs error ~1e-6 there): the reference evaluates sin(2^k x) in fp32, whose argument rounding alone is
// 2^k ulp(x), and the results are rounded to fp16 operands (5e-4), so the accurate sincosf (with its Payne-Hanek slow path:
// ~100 instructions, a constant-table walk) buys nothing.  |x| beyond ~1e5 (NDC of points far outside view 0) loses the
// reduction's accuracy -- exactly where the reference's own fp32 sin(512 x) is noise.
__device__ __forceinline__ void sincos_reduced(float x, float& s, float& c) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);          // 2 pi = hi + lo
  r = fmaf(n, 1.7484555e-7f, r);
  s = __sinf(r);
  c = __cosf(r);
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// ---- ray transformer for the 32 rows of one warp (ray_transformer.py:49-79, cond_nerf.py:83-99), on mma.sync.
// Phase A: [q | k | v] = raw . Wqkv^T (6 n-tiles per 16-row m-tile); q stays in registers as the A fragment of the
// score product, k and v go to shared memory as fp16 B operands.  Phase B, per m-tile and head: scores = Q_h . K^T
// with the other heads' columns of Q zeroed (so all four heads share the same K fragments, K = 16 = 4 heads x 4 dims),
// row maxima (pass 1), then the same product again with the accumulator initialised to -max, exp2, and P . V_h where
// V_h carries a column of ones: the softmax denominator drops out of the same MMA.  fc (+ residual), LayerNorm,
// out_alpha_linear.0 are two more MMAs on C-layout registers; sigma is written to sm.sig for the compositing scan.
template <int kAct, int kS>
__device__ __forceinline__ void ray_transformer_mma(TcSmem& sm, const int slot, const uint32_t buf, const int quarter, const int lane,
                                                    [[maybe_unused]] const unsigned it, [[maybe_unused]] unsigned& trace_n) {
  const int g = lane >> 2, t = lane & 3;
  const bool lo = t < 2;
  const float* hand = &sm.hand[slot][buf][0][0];
  const float* hand_nv = &sm.hand_nv[slot][buf][0];
  // Both m-tile loops are ROLLED (#pragma unroll 1): the kernel's instruction footprint, not its instruction count, is what
  // the timeline showed to be expensive -- every phase of every role used to run from a cold instruction cache (~90 KB of
  // straight-line code per tile against a 32 KB L1.5 I-cache).  Fragments that outlive the loop are kept with selects.
  uint32_t qa0[4] = {0u, 0u, 0u, 0u}, qa1[4] = {0u, 0u, 0u, 0u};   // Q as A fragments (fp16 pairs), per m-tile
  // ---------------- phase A
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = quarter * 32 + mt * 16 + g;     // fragment rows r0 and r0 + 8
    const float2 x00 = *reinterpret_cast<const float2*>(hand + r0 * kHandRow + 2 * t);
    const float2 x10 = *reinterpret_cast<const float2*>(hand + (r0 + 8) * kHandRow + 2 * t);
    const float2 x01 = *reinterpret_cast<const float2*>(hand + r0 * kHandRow + 8 + 2 * t);
    const float2 x11 = *reinterpret_cast<const float2*>(hand + (r0 + 8) * kHandRow + 8 + 2 * t);
    const uint32_t a0 = pack_h2(x00.x, x00.y), a1 = pack_h2(x10.x, x10.y), a2 = pack_h2(x01.x, x01.y), a3 = pack_h2(x11.x, x11.y);
    // x = hi + lo in fp16 pairs: scores reach ~1e3 (exp2 domain), so q and k want more than one fp16 of the inputs
    const uint32_t l0 = pack_h2(x00.x - h2_lo(a0), x00.y - h2_hi(a0)), l1 = pack_h2(x10.x - h2_lo(a1), x10.y - h2_hi(a1));
    const uint32_t l2 = pack_h2(x01.x - h2_lo(a2), x01.y - h2_hi(a2)), l3 = pack_h2(x11.x - h2_lo(a3), x11.y - h2_hi(a3));
    // cond_nerf.py:83: the mask disables whole QUERY rows (all scores equal -> uniform attention): zero their q
    const bool v0 = hand_nv[r0] > 1.f, v1 = hand_nv[r0 + 8] > 1.f;
    float c[6][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.p.wqkv_h[8 * j + g][2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.p.wqkv_h[8 * j + g][8 + 2 * t]);
      const uint32_t e0 = *reinterpret_cast<const uint32_t*>(&sm.p.wqkv_l[8 * j + g][2 * t]);
      const uint32_t e1 = *reinterpret_cast<const uint32_t*>(&sm.p.wqkv_l[8 * j + g][8 + 2 * t]);
      c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
      mma16816(c[j], a0, a1, a2, a3, e0, e1);
      mma16816(c[j], l0, l1, l2, l3, b0, b1);
      mma16816(c[j], a0, a1, a2, a3, b0, b1);
    }
    {
      const uint32_t q[4] = {v0 ? pack_h2(c[0][0], c[0][1]) : 0u, v1 ? pack_h2(c[0][2], c[0][3]) : 0u,
                             v0 ? pack_h2(c[1][0], c[1][1]) : 0u, v1 ? pack_h2(c[1][2], c[1][3]) : 0u};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qa0[i] = mt == 0 ? q[i] : qa0[i];
        qa1[i] = mt == 0 ? qa1[i] : q[i];
      }
    }
    __half* k0 = &sm.kbuf[slot][r0][0];
    __half* k1 = &sm.kbuf[slot][r0 + 8][0];
    *reinterpret_cast<uint32_t*>(k0 + 2 * t) = pack_h2(c[2][0], c[2][1]);
    *reinterpret_cast<uint32_t*>(k1 + 2 * t) = pack_h2(c[2][2], c[2][3]);
    *reinterpret_cast<uint32_t*>(k0 + 8 + 2 * t) = pack_h2(c[3][0], c[3][1]);
    *reinterpret_cast<uint32_t*>(k1 + 8 + 2 * t) = pack_h2(c[3][2], c[3][3]);
    // value columns: n-tile 4 = heads 0,1, n-tile 5 = heads 2,3; thread t holds dims 2(t&1).. of head (t>>1) (+2)
    const int vcol = (t >> 1) * 8 + (t >> 1) * 4 + (t & 1) * 2;   // odd heads keep their dims in columns 4..7
    __half* v0p = &sm.vbuf[slot][r0][vcol];
    __half* v1p = &sm.vbuf[slot][r0 + 8][vcol];
    *reinterpret_cast<uint32_t*>(v0p) = pack_h2(c[4][0], c[4][1]);
    *reinterpret_cast<uint32_t*>(v1p) = pack_h2(c[4][2], c[4][3]);
    *reinterpret_cast<uint32_t*>(v0p + 16) = pack_h2(c[5][0], c[5][1]);
    *reinterpret_cast<uint32_t*>(v1p + 16) = pack_h2(c[5][2], c[5][3]);
  }
  TRACE_RAY(2);
  if constexpr (kS == 64) ray_pair_barrier(slot, quarter);   // a ray's keys / values come from two warps,
  else ray_barrier_s<(kS > 128)>(slot);                       // ... from up to four (eight when S = 256)

  TRACE_RAY(3);
  // ---------------- phase B
  const float2 lnw0 = *reinterpret_cast<const float2*>(&sm.p.ln_w[2 * t]), lnw1 = *reinterpret_cast<const float2*>(&sm.p.ln_w[8 + 2 * t]);
  const float2 lnb0 = *reinterpret_cast<const float2*>(&sm.p.ln_b[2 * t]), lnb1 = *reinterpret_cast<const float2*>(&sm.p.ln_b[8 + 2 * t]);
#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int rbase = quarter * 32 + mt * 16;      // first row of the m-tile; it lies inside one ray (kS >= 16)
    uint32_t qm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) qm[i] = mt == 0 ? qa0[i] : qa1[i];
    // first key row of that ray; S = 256: the ray's keys are the 256 rows of kbuf[0] followed by kbuf[1] (contiguous in memory)
    const int key0 = kS > 128 ? 0 : (rbase & ~(kS - 1));
    const int kslot = kS > 128 ? 0 : slot;
    // ldmatrix lane addresses.  K (x4): matrices (n-tile 2np, k 0-7), (2np, k 8-15), (2np+1, k 0-7), (2np+1, k 8-15).
    const uint32_t kaddr = tc::smem_u32(&sm.kbuf[kslot][key0 + (lane >> 4) * 8 + (lane & 7)][((lane >> 3) & 1) * 8]);
    // V (x2, transposed): matrices (keys 16np .. +7), (keys 16np + 8 .. +15), 8 columns of one head
    const uint32_t vaddr = tc::smem_u32(&sm.vbuf[kslot][key0 + ((lane >> 3) & 1) * 8 + (lane & 7)][0]);
    uint32_t att[4] = {0u, 0u, 0u, 0u};            // normalised attention output as the A fragment of fc
#pragma unroll 1
    for (int hp = 0; hp < 2; ++hp) {               // head pair (0,1) / (2,3)
      float o[2][4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * hp + hh;
        // Q_h: only the 4 k-columns of head h survive (k 0-3: t<2 of a0/a1; 4-7: t>=2; 8-11: t<2 of a2/a3; 12-15: t>=2)
        const bool mine = (hh == 0) ? lo : !lo;
        const uint32_t q0 = mine ? (hp == 0 ? qm[0] : qm[2]) : 0u, q1 = mine ? (hp == 0 ? qm[1] : qm[3]) : 0u;
        // The keys are walked in blocks of <= 64.  Inside a block the head's scores of this m-tile (16 rows x 64 keys = 32 registers
        // per lane) stay in registers between the row-maximum pass and the exponentials -- ONE score product per key instead of the two
        // of a two-pass softmax -- and blocks are merged flash-attention style (exact: both partial results are rescaled to the common
        // maximum; the softmax denominator rides in the ones-column of V and is rescaled with it).  The legacy mma.sync pipe is what
        // bounds the ray groups on this part (~27 cycles per m16n8k16 per SM sub-partition; ncu: `math pipe throttle` is their top stall).
        constexpr int kBlk = kS < 64 ? kS : 64;
        float m0 = -INFINITY, m1 = -INFINITY;      // running maxima of rows g, g + 8
        o[hh][0] = o[hh][1] = o[hh][2] = o[hh][3] = 0.f;
        auto do_block = [&](const int blk) {
          const uint32_t kblk = kaddr + blk * kBlk * kWRow * 2, vblk = vaddr + blk * kBlk * kVRow * 2 + h * 16;
          float sc[kBlk / 16][8];
          float b0 = -INFINITY, b1 = -INFINITY;
#pragma unroll
          for (int np = 0; np < kBlk / 16; ++np) {
            uint32_t kf[4];
            ldsm_x4(kf, kblk + np * 16 * kWRow * 2);
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
            // K = 8 products: only the k-half that holds this head pair's dimensions is multiplied (the other half of Q_h is zero)
            mma1688(c0, q0, q1, hp == 0 ? kf[0] : kf[1]);
            mma1688(c1, q0, q1, hp == 0 ? kf[2] : kf[3]);
            b0 = fmax3(b0, c0[0], c0[1]); b0 = fmax3(b0, c1[0], c1[1]);
            b1 = fmax3(b1, c0[2], c0[3]); b1 = fmax3(b1, c1[2], c1[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sc[np][i] = c0[i]; sc[np][4 + i] = c1[i]; }
          }
          b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 1)); b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, 2));
          b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 1)); b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, 2));
          float ob[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int np = 0; np < kBlk / 16; ++np) {
            const uint32_t p0 = pack_h2(ex2_fast(sc[np][0] - b0), ex2_fast(sc[np][1] - b0)), p1 = pack_h2(ex2_fast(sc[np][2] - b1), ex2_fast(sc[np][3] - b1));
            const uint32_t p2 = pack_h2(ex2_fast(sc[np][4] - b0), ex2_fast(sc[np][5] - b0)), p3 = pack_h2(ex2_fast(sc[np][6] - b1), ex2_fast(sc[np][7] - b1));
            uint32_t vf0, vf1;
            ldsm_x2_trans(vf0, vf1, vblk + np * 16 * kVRow * 2);
            mma16816(ob, p0, p1, p2, p3, vf0, vf1);
          }
          if constexpr (kS / kBlk == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[hh][i] = ob[i];
          } else {
            const float n0 = fmaxf(m0, b0), n1 = fmaxf(m1, b1);
            const float r0 = ex2_fast(m0 - n0), r1 = ex2_fast(m1 - n1), s0 = ex2_fast(b0 - n0), s1 = ex2_fast(b1 - n1);   // first block: ex2(-inf) = 0
            o[hh][0] = o[hh][0] * r0 + ob[0] * s0; o[hh][1] = o[hh][1] * r0 + ob[1] * s0;
            o[hh][2] = o[hh][2] * r1 + ob[2] * s1; o[hh][3] = o[hh][3] * r1 + ob[3] * s1;
            m0 = n0; m1 = n1;
          }
        };
        if constexpr (kS / kBlk == 1) {
          do_block(0);
        } else {
#pragma unroll 1
          for (int blk = 0; blk < kS / kBlk; ++blk) do_block(blk);
        }
      }
      // even head: (o0,o1 | o2,o3 | den,0 | 0,0) over t = 0..3; odd head: (den,0 | 0,0 | o0,o1 | o2,o3).  fc wants
      // k = head*4 + dim, i.e. exactly the even head's values for t < 2 and the odd head's for t >= 2.
      const int src = (lane & ~3) + (lo ? 2 : 0);
      const float den0 = __shfl_sync(0xffffffffu, t == 0 ? o[1][0] : o[0][0], src);
      const float den1 = __shfl_sync(0xffffffffu, t == 0 ? o[1][2] : o[0][2], src);
      const float i0 = rcp_fast(den0), i1 = rcp_fast(den1);
      {
        const uint32_t e0 = pack_h2((lo ? o[0][0] : o[1][0]) * i0, (lo ? o[0][1] : o[1][1]) * i0);
        const uint32_t e1 = pack_h2((lo ? o[0][2] : o[1][2]) * i1, (lo ? o[0][3] : o[1][3]) * i1);
        att[0] = hp == 0 ? e0 : att[0]; att[1] = hp == 0 ? e1 : att[1];
        att[2] = hp == 0 ? att[2] : e0; att[3] = hp == 0 ? att[3] : e1;
      }
    }
    TRACE_RAY(4);
    // fc + residual (C layout: y[j] = columns 8j + 2t, 8j + 2t + 1 of rows g | g + 8)
    const int r0 = rbase + g;
    float y[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 xa = *reinterpret_cast<const float2*>(hand + r0 * kHandRow + 8 * j + 2 * t);
      const float2 xb = *reinterpret_cast<const float2*>(hand + (r0 + 8) * kHandRow + 8 * j + 2 * t);
      y[j][0] = xa.x; y[j][1] = xa.y; y[j][2] = xb.x; y[j][3] = xb.y;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.p.fc_h[8 * j + g][2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.p.fc_h[8 * j + g][8 + 2 * t]);
      mma16816(y[j], att[0], att[1], att[2], att[3], b0, b1);
    }
    // LayerNorm(eps 1e-6) over the 16 columns: 4 per thread, the quad holds the row
    float s0 = (y[0][0] + y[0][1]) + (y[1][0] + y[1][1]), s1 = (y[0][2] + y[0][3]) + (y[1][2] + y[1][3]);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float mu0 = s0 * (1.f / 16.f), mu1 = s1 * (1.f / 16.f);
#pragma unroll
    for (int j = 0; j < 2; ++j) { y[j][0] -= mu0; y[j][1] -= mu0; y[j][2] -= mu1; y[j][3] -= mu1; }
    float q0 = y[0][0] * y[0][0] + y[0][1] * y[0][1] + y[1][0] * y[1][0] + y[1][1] * y[1][1];
    float q1 = y[0][2] * y[0][2] + y[0][3] * y[0][3] + y[1][2] * y[1][2] + y[1][3] * y[1][3];
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float rs0 = rsqrtf(q0 * (1.f / 16.f) + 1e-6f), rs1 = rsqrtf(q1 * (1.f / 16.f) + 1e-6f);
    const uint32_t n0 = pack_h2(fmaf(y[0][0] * rs0, lnw0.x, lnb0.x), fmaf(y[0][1] * rs0, lnw0.y, lnb0.y));
    const uint32_t n1 = pack_h2(fmaf(y[0][2] * rs1, lnw0.x, lnb0.x), fmaf(y[0][3] * rs1, lnw0.y, lnb0.y));
    const uint32_t n2 = pack_h2(fmaf(y[1][0] * rs0, lnw1.x, lnb1.x), fmaf(y[1][1] * rs0, lnw1.y, lnb1.y));
    const uint32_t n3 = pack_h2(fmaf(y[1][2] * rs1, lnw1.x, lnb1.x), fmaf(y[1][3] * rs1, lnw1.y, lnb1.y));
    // out_alpha_linear: 16 -> 16 (MMA, bias as the initial accumulator), activation, 16 -> 1
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float2 bj = *reinterpret_cast<const float2*>(&sm.p.oa0_b[8 * j + 2 * t]);
      const float2 wj = *reinterpret_cast<const float2*>(&sm.p.oa2_w[8 * j + 2 * t]);
      float c[4] = {bj.x, bj.y, bj.x, bj.y};
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sm.p.oa0_h[8 * j + g][2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sm.p.oa0_h[8 * j + g][8 + 2 * t]);
      mma16816(c, n0, n1, n2, n3, b0, b1);
      d0 = fmaf(act_fn<kAct>(c[0]), wj.x, fmaf(act_fn<kAct>(c[1]), wj.y, d0));
      d1 = fmaf(act_fn<kAct>(c[2]), wj.x, fmaf(act_fn<kAct>(c[3]), wj.y, d1));
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    TRACE_RAY(8);
    if (t == 0) {
      sm.sig[slot][r0] = fmaxf(d0 + sm.p.oa2_b, 0.f);
      sm.sig[slot][r0 + 8] = fmaxf(d1 + sm.p.oa2_b, 0.f);
    }
  }
  __syncwarp();
}


// ---- alpha compositing of the rows of one warp (nerf.py:101-124 with wo_render_interval: alpha = 1 - exp(-sigma),
// T_i = exp(-sum_{j<i} sigma_j), out = sum_i T_i alpha_i x_i for x = r, g, b, depth, 1).
// Each warp composites ITS segment of the ray with a transmittance that starts at 1 (shuffle scan + a butterfly that reduces the five
// sums with 9 shuffles instead of 25), and publishes (five partial sums, its optical depth).  Segments combine as
// out = sum_w exp(-sum_{w' < w} tau_w') part_w, evaluated by the ray's first lane after ONE barrier per tile (the same barrier also
// orders this tile's key / value reads before the next tile's writes; `red` is double-buffered by tile parity).
// Returns true in the lane that holds the ray's result.
template <int kS>
__device__ __forceinline__ bool composite_rows(TcSmem& sm, const int slot, const int quarter, const int lane, const unsigned it,
                                               const float sigma, const bool valid, const float (&x)[5], float (&out)[5]) {
  constexpr int kSeg = kS < 32 ? kS : 32;          // rows of one ray inside a warp
  constexpr int kWarps = kS < 32 ? 1 : kS / 32;    // warps per ray
  float incl = sigma;
#pragma unroll
  for (int off = 1; off < kSeg; off <<= 1) {
    const float nb = __shfl_up_sync(0xffffffffu, incl, off, 32);
    if ((lane & (kSeg - 1)) >= off) incl += nb;
  }
  const float wgt = valid ? __expf(sigma - incl) * (1.f - __expf(-sigma)) : 0.f;
  float v[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) v[i] = wgt * x[i];
  if constexpr (kSeg < 32) {
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int off = kSeg / 2; off >= 1; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
#pragma unroll
    for (int i = 0; i < 5; ++i) out[i] = v[i];
    return (lane & (kSeg - 1)) == 0;
  } else {
    // butterfly: after the xor-16 / 8 / 4 steps lane l holds value index ((l >> 4) & 1) * 4 + ((l >> 3) & 1) * 2 + ((l >> 2) & 1)
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a4[4];
    a4[0] = (b4 ? v[4] : v[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v[0] : v[4], 16);
#pragma unroll
    for (int i = 1; i < 4; ++i) a4[i] = (b4 ? 0.f : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : 0.f, 16);
    float a2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) a2[i] = (b3 ? a4[i + 2] : a4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? a4[i] : a4[i + 2], 8);
    float a1 = (b2 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? a2[0] : a2[1], 4);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    const float tau = __shfl_sync(0xffffffffu, incl, 31);        // optical depth of this warp's segment
    if constexpr (kWarps == 1) {
#pragma unroll
      for (int i = 0; i < 5; ++i) out[i] = __shfl_sync(0xffffffffu, a1, (i & 4) * 4 + (i & 2) * 4 + (i & 1) * 4);
      return lane == 0;
    } else {
      constexpr bool kBoth = kS > 128;                            // S = 256: the ray spans both ray groups
      const int wq = kBoth ? slot * 4 + quarter : quarter;        // warp index inside the scan domain
      float (*red)[8] = kBoth ? &sm.red[it & 1][0][0] : &sm.red[it & 1][slot][0];
      if ((lane & 3) == 0 && lane < 20) red[wq][(lane >> 4) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = a1;
      if (lane == 31) red[wq][5] = tau;
      if constexpr (kBoth) ray_barrier_s<true>(slot);
      else if constexpr (kS == 64) ray_pair_barrier(slot, quarter);
      else ray_barrier(slot);
      const int w0 = wq & ~(kWarps - 1);                          // first warp of this ray
      const bool writer = lane == 0 && wq == w0;
      if (writer) {
        float T = 1.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) out[i] = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          const float4 p = *reinterpret_cast<const float4*>(&red[w0 + w][0]);
          const float2 q = *reinterpret_cast<const float2*>(&red[w0 + w][4]);
          out[0] = fmaf(T, p.x, out[0]); out[1] = fmaf(T, p.y, out[1]); out[2] = fmaf(T, p.z, out[2]); out[3] = fmaf(T, p.w, out[3]);
          out[4] = fmaf(T, q.x, out[4]);
          T *= __expf(-q.y);
        }
      }
      return writer;
    }
  }
}

}  // namespace

struct DecoderWeightsTC {
  unsigned char* packed = nullptr;   // device: kPackedBytes of pre-swizzled fp16 chunks
  TcParams* params = nullptr;        // device
  float4* posenc = nullptr;          // device: [kMaxSamples][16] sinusoid table of the ray transformer (cond_nerf.py:118-127)
};

// ------------------------------------------------------------------------------------------------------------------
template <int kAct>
__global__ void __launch_bounds__(kThreads, 1)
decoder_tc_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const mnf_decoder_cfg cfg,
                  const unsigned char* __restrict__ wpacked, const TcParams* __restrict__ gparams,
                  const float4* __restrict__ posenc_tab, const __half* __restrict__ cond, const int setbg_opaque, float* __restrict__ out_rgb,
                  float* __restrict__ out_depth, float* __restrict__ out_opacity, float* __restrict__ aux) {
  extern __shared__ unsigned char smem_dyn[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = cfg.n_samples;
  // S <= 128: a 128-sample tile holds 128 / S whole rays.  S = 256: a tile is HALF a ray and the two slots of a pair hold one ray.
  const bool split_ray = S > kTileM;
  const int rays_per_tile = split_ray ? 1 : kTileM / S;
  const int64_t n_tiles = split_ray ? 2 * rays.n_rays : (rays.n_rays + rays_per_tile - 1) / rays_per_tile;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  // ---- one-time setup
  for (int i = tid; i < (int)(sizeof(TcParams) / 4); i += blockDim.x)
    reinterpret_cast<float*>(&sm.p)[i] = reinterpret_cast<const float*>(gparams)[i];
  for (int i = tid; i < 2 * kTileM * kVRow; i += blockDim.x) {   // the constant columns of the value operand (see vbuf)
    const int c = i % kVRow, head = c >> 3, e = c & 7;
    const bool one = head < 4 && ((head & 1) ? e == 0 : e == 4);
    (&sm.vbuf[0][0][0])[i] = __float2half_rn(one ? 1.f : 0.f);
  }
  if (tid == 0) {
    sm.trace.buf = g_trace_buf;
    sm.trace.per = (g_trace_buf != nullptr && blockIdx.x == 0) ? g_trace_cap / 20u : 0u;
    for (int i = 0; i < kNumStages; ++i) {
      tc::mbar_init(&sm.w_full[i], 1);
      tc::mbar_init(&sm.w_empty[i], 1);
    }
    tc::mbar_init(&sm.bias_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sm.a_ready[i], kTileM);
      tc::mbar_init(&sm.d_full[i], 1);
      for (int j = 0; j < 2; ++j) {
        tc::mbar_init(&sm.ray_full[i][j], kTileM);
        tc::mbar_init(&sm.ray_empty[i][j], kTileM);
        tc::mbar_init(&sm.geo_full[i][j], 32);
        tc::mbar_init(&sm.geo_empty[i][j], kTileM);
      }
    }
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const int wg = warp >> 2;      // warpgroup: 0 control, 1-2 trunk slots, 3-4 ray groups

  if (wg == 0) {
    reg_dec<kRegsCtrl>();
    if (warp == 0) {
      // ================================================================== weight streamer
      if (lane == 0) {
        tc::mbar_arrive_expect_tx(&sm.bias_full, kChunkBytes);
        tc::bulk_g2s(sm.biasw, wpacked + kBiasOffset, kChunkBytes, &sm.bias_full);
        uint32_t n = 0;
        for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
          for (int c = 0; c < kNumChunks; ++c, ++n) {
            const uint32_t st = n % kNumStages, par = (n / kNumStages) & 1;
            mbar_wait_sleep(&sm.w_empty[st], par ^ 1, 200);
            tc::mbar_arrive_expect_tx(&sm.w_full[st], chunk_bytes(c));
            tc::bulk_g2s(sm.ring[st], wpacked + chunk_offset(c), chunk_bytes(c), &sm.w_full[st]);
          }
        }
      }
    } else if (warp == 1) {
      // ================================================================== MMA issuer
      // The tensor pipe idles whenever this warp is between "operand ready" and the next tcgen05.mma, so its loop is
      // kept lean: the 8 phases are unrolled (K-step structure is compile-time), a shared-memory descriptor is
      // (constant high word | low word), low word = ring base + stage * 1024 + K-step * 2 (units of 16 B), stages and
      // barrier parities are tracked incrementally (no div/mod by 6).  The whole warp runs the loop convergently and one
      // elected lane issues the tcgen05 instructions.
      const uint32_t idesc128 = tc::umma_idesc_f16(128, 128), idesc_head = tc::umma_idesc_f16(128, kHeadN);
      const bool leader = tc::elect_one();
      constexpr uint64_t kDescHi = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16);
      const uint32_t ring_lo = (tc::smem_u32(sm.ring[0]) >> 4) & 0x3FFFu;
      const uint32_t bias_lo = (tc::smem_u32(sm.biasw) >> 4) & 0x3FFFu;
      auto bdesc = [](uint32_t lo) { return kDescHi | (uint64_t)lo; };
      unsigned trace_m = 0u;
      uint32_t n = 0;            // chunk counter (trace only)
      uint32_t st = 0;           // ring stage of the next chunk
      uint32_t full_par = 0;     // bit s = parity of the next w_full[s] phase to wait for
      mbar_wait_sleep(&sm.bias_full, 0, 20);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const bool active1 = 2 * pair + 1 < n_tiles;
#pragma unroll
        for (int ph = 0; ph < kNumPhases; ++ph) {
          const int nch = ph <= 1 ? 1 : (ph == 6 ? 3 : 2);
          uint32_t stage[3], lo[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < nch) {
              stage[j] = st;
              st = (st == kNumStages - 1) ? 0u : st + 1u;
            } else {
              stage[j] = stage[0];
            }
            lo[j] = ring_lo + stage[j] * (uint32_t)(kChunkBytes >> 4);
          }
          // layer 5 (three chunks, a 4-stage ring): only its first chunk is awaited here, the other two right before the K steps
          // that read them, so that their copies run underneath the first four MMAs
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < nch && !(ph == 6 && j > 0)) {
              mbar_wait_sleep(&sm.w_full[stage[j]], (full_par >> stage[j]) & 1u, 20);
              full_par ^= 1u << stage[j];
            }
          TRACE_MMA(0, 50 + ph);
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            if (slot == 1 && !active1) continue;
            mbar_wait_sleep(&sm.a_ready[slot], ph & 1, 20);
            TRACE_MMA(slot, 10 + ph);
            tc::tc_fence_after_sync();
            uint32_t tb = tmem + slot * kSlotCols;
            asm volatile("" : "+r"(tb));   // opaque: keeps the 20-odd operand addresses of a phase from being hoisted into registers
            const uint32_t d = tb + kColD;
            if (leader) {
              if (ph == 0) {          // gate = pts_bias(cond): K = 32
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) tc::umma_ts(d, tb + kColCond + ks * 8, bdesc(lo[0] + ks * 2), idesc128, ks > 0);
              } else if (ph == 1) {   // layer 0: K = 64 (encoding)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc::umma_ts(d, tb + kColEnc + ks * 8, bdesc(lo[0] + ks * 2), idesc128, ks > 0);
              } else if (ph <= 5) {   // layers 1..4: K = 128, then the bias step (A = encoding columns 48..63, see biasw)
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(lo[ks >> 2] + (ks & 3) * 2), idesc128, ks > 0);
                tc::umma_ts(d, tb + kColEnc + 3 * 8, bdesc(bias_lo + (ph - 2) * 2), idesc128, 1);
              } else if (ph == 6) {   // layer 5 on [enc, h]: K = 64 + 128; the h part follows below (after its chunks have landed)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc::umma_ts(d, tb + kColEnc + ks * 8, bdesc(lo[0] + ks * 2), idesc128, ks > 0);
              } else {                // heads: [alpha_linear | views(feature_linear(.))], N = 80, K = 128
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(lo[ks >> 2] + (ks & 3) * 2), idesc_head, ks > 0);
              }
              if (ph != 6) tc::umma_commit(&sm.d_full[slot]);
            }
            if (ph == 6) {
              if (slot == 0) {
#pragma unroll
                for (int j = 1; j < 3; ++j) {
                  mbar_wait_sleep(&sm.w_full[stage[j]], (full_par >> stage[j]) & 1u, 20);
                  full_par ^= 1u << stage[j];
                }
              }
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) tc::umma_ts(d, tb + kColH + ks * 8, bdesc(lo[1 + (ks >> 2)] + (ks & 3) * 2), idesc128, 1);
                tc::umma_commit(&sm.d_full[slot]);
              }
            }
            TRACE_MMA(slot, 30 + ph);
            __syncwarp();
          }
          if (leader) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (j < nch) tc::umma_commit(&sm.w_empty[stage[j]]);
          }
          TRACE_MMA(0, 60 + ph);
          __syncwarp();
          n += nch;
        }
      }
      (void)n; (void)trace_m;
    } else {
      // ================================================================== ray geometry (warp 2 -> trunk slot 0, warp 3 -> slot 1)
      // Everything a tile needs PER RAY is computed here, once, one tile ahead of the trunk slot: the ray direction (misc/camera.py:
      // 255-278), the slope of the view-0 projection along the ray, q(t) = K0 (R0 (o + t d) + T0) = q_o + t K0 R0 d (camera.py:351-379 is
      // linear in the depth before its perspective division), and the direction term of the colour head (64 values per ray).  The
      // trunk threads used to derive all of it per SAMPLE -- ~250 instructions on the critical path of every tile -- and are left
      // with one fused multiply-add per coordinate.
      const int slot = warp - 2;
      const float* E = cams.w2c[0];
      const float* K = cams.K[0];
      uint32_t n = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++n) {
        const int64_t tile = 2 * pair + slot;
        if (tile >= n_tiles) break;
        const uint32_t nb = n & 1u;
        if (n >= 2) mbar_wait_sleep(&sm.geo_empty[slot][nb], ((n >> 1) - 1u) & 1u, 64);
        float dir[3] = {0.f, 0.f, 0.f};
        if (lane < rays_per_tile) {
          const int64_t ray_n = split_ray ? pair : tile * rays_per_tile + lane;
          float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ray_n < rays.n_rays) {
            const int64_t pix = rays.ray_idx ? rays.ray_idx[ray_n] : rays.first_ray + ray_n;
            float o[3], d[3];
            cast_ray(cams, pix, o, d);
            float rd[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) rd[i] = fmaf(d[2], E[i * 4 + 2], fmaf(d[1], E[i * 4 + 1], d[0] * E[i * 4 + 0]));
            rec.x = fmaf(rd[2], K[2], fmaf(rd[1], K[1], rd[0] * K[0]));
            rec.y = fmaf(rd[2], K[5], fmaf(rd[1], K[4], rd[0] * K[3]));
            rec.z = fmaf(rd[2], K[8], fmaf(rd[1], K[7], rd[0] * K[6]));
            rec.w = 1.f;
            const float inv = rsqrtf(fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-24f));
#pragma unroll
            for (int i = 0; i < 3; ++i) dir[i] = rd[i] * inv;                  // unit direction in the frame of source view 0 (matchnerf.py:129-134)
          }
          sm.rayrec[slot][nb][lane] = rec;
        }
        for (int rl = 0; rl < rays_per_tile; ++rl) {
          const float dx = __shfl_sync(0xffffffffu, dir[0], rl), dy = __shfl_sync(0xffffffffu, dir[1], rl), dz = __shfl_sync(0xffffffffu, dir[2], rl);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int o2 = lane + 32 * h;
            sm.dirvec[slot][nb][rl][o2] = sm.p.views_dir[o2 * 3] * dx + sm.p.views_dir[o2 * 3 + 1] * dy + sm.p.views_dir[o2 * 3 + 2] * dz + sm.p.views_b[o2];
          }
        }
        tc::mbar_arrive(&sm.geo_full[slot][nb]);             // every lane releases its own writes (count 32)
      }
    }
  } else if (wg <= 2) {
    // ================================================================== trunk slot: staging, epilogues, heads
    reg_inc<kRegsTrunk>();
    const int slot = wg - 1;
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;             // sample row inside the tile == TMEM lane
    const uint32_t tb = tmem + slot * kSlotCols + ((uint32_t)(quarter * 32) << 16);
    const int ray_local = split_ray ? 0 : row / S, s = split_ray ? slot * kTileM + row : row - ray_local * S;
    uint32_t it = 0;
    unsigned trace_n = 0;
    uint32_t* const srow = &sm.stage[slot][row][0];

    // ---- staging of a sample = its A operands (positional encoding, conditioning row) + depth + the ray's direction term.
    // It is SOFTWARE-PIPELINED across tiles: the timeline (tools/decoder_trace.py) showed ~5,000 cycles per tile pair in which
    // both slots staged and the tensor pipe idled.  Now the sample of tile i + 1 is staged inside the windows in which tile i waits
    // for its accumulators (part A after the gate epilogue, part B after layer 0), into a shared-memory row the SAME thread reads
    // back at the start of the next tile; the conditioning row travels by cp.async, so it occupies no registers in between.
    // q_o = K0 (R0 o + T0): the view-0 projection of the target camera centre (every ray starts there)
    float qo[3];
    {
      const float* E = cams.w2c[0];
      const float* K = cams.K[0];
      const float o[3] = {cams.c2w[3], cams.c2w[7], cams.c2w[11]};
      float c[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) c[i] = fmaf(o[2], E[i * 4 + 2], fmaf(o[1], E[i * 4 + 1], fmaf(o[0], E[i * 4 + 0], E[i * 4 + 3])));
#pragma unroll
      for (int i = 0; i < 3; ++i) qo[i] = fmaf(c[2], K[i * 3 + 2], fmaf(c[1], K[i * 3 + 1], c[0] * K[i * 3 + 0]));
    }
    const float rW1 = __frcp_rn((float)(cams.W - 1)), rH1 = __frcp_rn((float)(cams.H - 1));
    const float rS1 = (cams.tfar - cams.tnear) * __frcp_rn((float)(S - 1)), rNF = __frcp_rn(cams.nf[0][1] - cams.nf[0][0]);
    // `n_it` = index of the tile among this CTA's tiles (selects the parity of its ray-record buffer)
    auto stage_geometry = [&](const int64_t pair_n, const uint32_t n_it, float (&x)[3]) {
      const uint32_t nb = n_it & 1u;
      const int64_t tile_n = 2 * pair_n + slot;
      const int64_t ray_n = split_ray ? pair_n : tile_n * rays_per_tile + ray_local;
      const bool valid_n = ray_n < rays.n_rays;
      const size_t n_glob_n = valid_n ? (size_t)ray_n * S + s : 0;
      x[0] = x[1] = x[2] = 0.f;
      float depth_n = 0.f;
      if (valid_n) {
        const uint32_t dst = tc::smem_u32(srow + 32);
        const char* src = reinterpret_cast<const char*>(cond + n_glob_n * kCondPad);
#pragma unroll
        for (int j = 0; j < 4; ++j) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * j), "l"(src + 16 * j) : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(srow + 32 + 4 * j) = make_uint4(0u, 0u, 0u, 0u);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      TRACE_TRUNK(30);
      mbar_wait_sleep(&sm.geo_full[slot][nb], (n_it >> 1) & 1u, 32);      // the tile's ray records and dirvec (written a tile ahead)
      TRACE_TRUNK(31);
      if (valid_n) {
        const float4 rec = sm.rayrec[slot][nb][ray_local];
        const float u = rays.jitter ? rays.jitter[n_glob_n] : 0.f;
        // matchnerf.py:163-181 (legacy): t = near + (i + u) / (S - 1) (far - near); the decoder's copy only feeds fp16 operands and the
        // depth output, so the reciprocal-multiply form (<= 1 ulp) replaces the IEEE division the gather keeps for its mask decisions
        depth_n = fmaf((float)s + u, rS1, cams.tnear);
        const float q0 = fmaf(depth_n, rec.x, qo[0]), q1 = fmaf(depth_n, rec.y, qo[1]), q2 = fmaf(depth_n, rec.z, qo[2]);
        const float rz = rcp_fast(q2);                                  // feeds fp16 operands only
        x[0] = q0 * rz * rW1;
        x[1] = q1 * rz * rH1;
        x[2] = (q2 - cams.nf[0][0]) * rNF;
      }
      TRACE_TRUNK(34);
      srow[48] = __float_as_uint(depth_n);
    };
    auto stage_encoding = [&](const float (&x)[3]) {
      // encoding order (cond_nerf.py:108-116, :56-57): x, sin(2^k x) k-major, cos(2^k x) k-major, pad
      float sn[3], cs[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) sincos_reduced(x[i], sn[i], cs[i]);
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
      put(63, 1.f);   // pad column = 1: multiplies the bias column of the layer-0 / layer-5 weights
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(srow + 4 * j) = make_uint4(e[4 * j], e[4 * j + 1], e[4 * j + 2], e[4 * j + 3]);
    };
    // ---- a staged row -> fp16 A operands in tensor memory (positional encoding, conditioning), then "operands ready" to the MMA issuer.
    // Runs in the prologue for the CTA's first tile and, for every later tile, at the END of the previous tile right after the heads
    // accumulator has been read into registers: the gate MMA of tile i + 1 then runs underneath the colour-head arithmetic of tile i.
    auto stage_to_tmem = [&](float& depth_o, float& nvs_o) {
      asm volatile("cp.async.wait_all;" ::: "memory");
      // three 16-register groups, one after the other (the call at the end of a tile runs while the 80 head values are live)
      auto load16 = [&](uint32_t (&r)[16], const int word0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 a4 = *reinterpret_cast<const uint4*>(srow + word0 + 4 * j);
          r[4 * j] = a4.x; r[4 * j + 1] = a4.y; r[4 * j + 2] = a4.z; r[4 * j + 3] = a4.w;
        }
      };
      {
        uint32_t e[16];
        load16(e, 0);
        tc::tmem_st16(tb + kColEnc, e);
      }
      {
        uint32_t e[16];
        load16(e, 16);
        tc::tmem_st16(tb + kColEnc + 16, e);
      }
      {
        uint32_t cnd[16];
        load16(cnd, 32);
        // visibility masks live at cond[19..21]
        const __half2 h9 = *reinterpret_cast<const __half2*>(&cnd[9]);
        const __half2 h10 = *reinterpret_cast<const __half2*>(&cnd[10]);
        nvs_o = __high2float(h9) + __low2float(h10) + __high2float(h10);
        cnd[11] = (cnd[11] & 0xffff0000u) | 0x3c00u;   // pad column 22 = 1.0: multiplies the gate's bias column
        tc::tmem_st16(tb + kColCond, cnd);
      }
      depth_o = __uint_as_float(srow[48]);
      tc::tmem_wait_st();
      tc::tc_fence_before_sync();
      tc::mbar_arrive(&sm.a_ready[slot]);
    };
    float depth_t = 0.f, n_views_seen = 0.f;        // of the tile whose operands are in tensor memory
    {   // prologue: the first tile of this CTA is staged up front
      const int64_t pair0 = blockIdx.x;
      if (pair0 < n_pairs && 2 * pair0 + slot < n_tiles) {
        float x0[3];
        stage_geometry(pair0, 0u, x0);
        stage_encoding(x0);
        stage_to_tmem(depth_t, n_views_seen);
      }
    }

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const int64_t tile = 2 * pair + slot;
      if (tile >= n_tiles) break;
      const int64_t ray = split_ray ? pair : tile * rays_per_tile + ray_local;
      const bool valid = ray < rays.n_rays;
      const size_t n_glob = valid ? (size_t)ray * S + s : 0;
      const uint32_t cb = it & 1u;                                  // dirvec buffer of this tile
      const int64_t pair_next = pair + gridDim.x;
      const bool has_next = pair_next < n_pairs && 2 * pair_next + slot < n_tiles;
      (void)n_glob;

      TRACE_TRUNK(0);
      float xn[3] = {0.f, 0.f, 0.f};                                // NDC point of the NEXT tile's sample (part A -> part B)

      // ---------------- gate = pts_bias(cond) (bias folded into the MMA), kept as 64 packed-half registers for all six layers
      uint32_t gate[64];
      TRACE_TRUNK(1);
      mbar_wait_sleep(&sm.d_full[slot], 0, 32);
      TRACE_TRUNK(2);
      tc::tc_fence_after_sync();
      {   // 16-column chunks, the load of chunk c + 1 in flight while chunk c is converted (tcgen05.wait::ld covers all loads issued so far)
        uint32_t ra[16], rb[16];
        tc::tmem_ld16(tb + kColD, ra);
        tc::tmem_wait_ld(ra);
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          tc::tmem_ld16(tb + kColD + 16 * (c + 1), rb);
#pragma unroll
          for (int j = 0; j < 8; ++j) gate[8 * c + j] = pack_h2(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
          tc::tmem_wait_ld(rb);
          if (c + 2 < 8) tc::tmem_ld16(tb + kColD + 16 * (c + 2), ra);
#pragma unroll
          for (int j = 0; j < 8; ++j) gate[8 * (c + 1) + j] = pack_h2(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
          if (c + 2 < 8) tc::tmem_wait_ld(ra);
        }
      }
      tc::tc_fence_before_sync();
      tc::mbar_arrive(&sm.a_ready[slot]);

      // ---------------- trunk: h = relu(acc * gate), acc = W h + b from the tensor pipe, -> fp16 -> tensor memory (next layer's A operand)
#pragma unroll 1
      for (int l = 0; l < kDepth; ++l) {
        TRACE_TRUNK(10 + l);
        mbar_wait_sleep(&sm.d_full[slot], (l + 1) & 1, 32);
        TRACE_TRUNK(20 + l);
        tc::tc_fence_after_sync();
        {   // software-pipelined epilogue: 16 accumulator columns per step, the next step's tcgen05.ld issued before this step's math
          uint32_t ra[16], rb[16];
          tc::tmem_ld16(tb + kColD, ra);
          tc::tmem_wait_ld(ra);
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            uint32_t o8[8];
            tc::tmem_ld16(tb + kColD + 16 * (c + 1), rb);
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = gate_relu(pack_h2(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1])), gate[8 * c + j]);
            tc::tmem_st8(tb + kColH + 8 * c, o8);
            tc::tmem_wait_ld(rb);
            if (c + 2 < 8) tc::tmem_ld16(tb + kColD + 16 * (c + 2), ra);
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = gate_relu(pack_h2(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1])), gate[8 * (c + 1) + j]);
            tc::tmem_st8(tb + kColH + 8 * (c + 1), o8);
            if (c + 2 < 8) tc::tmem_wait_ld(ra);
          }
        }
        tc::tmem_wait_st();
        tc::tc_fence_before_sync();
        tc::mbar_arrive(&sm.a_ready[slot]);
        // the NEXT tile's sample is staged inside the two longest accumulator waits of this tile (dirvec[cb ^ 1] is free: every thread of
        // the slot finished the previous tile's heads arithmetic before this tile's layer-0 MMA could start)
        if (l == kStageGeomLayer && has_next) stage_geometry(pair_next, it + 1u, xn);
        if (l == kStageEncLayer && has_next) stage_encoding(xn);
      }

      // ---------------- heads: raw alpha (16) and the colour hidden layer (64)
      float xr[16];
      float rgb[3];
      TRACE_TRUNK(3);
      mbar_wait_sleep(&sm.d_full[slot], 1, 32);
      TRACE_TRUNK(4);
      tc::tc_fence_after_sync();
      const float depth_cur = depth_t, nvs_cur = n_views_seen;
      {
        // the whole heads accumulator (16 + 64 columns) moves to registers at once (the gate registers are dead by now), which frees
        // the accumulator columns: the NEXT tile's operands go to tensor memory right away and its gate MMA runs underneath the
        // colour-head arithmetic below
        uint32_t r16[16], rc0[32], rc1[32];
        tc::tmem_ld16(tb + kColD, r16);
        tc::tmem_ld32(tb + kColD + 16, rc0);
        tc::tmem_ld32(tb + kColD + 48, rc1);
        tc::tmem_wait_ld();
        if (has_next) stage_to_tmem(depth_t, n_views_seen);
#pragma unroll
        for (int j = 0; j < 16; ++j) xr[j] = act_fn<kAct>(__uint_as_float(r16[j]) + sm.p.alpha_b[j]);
        if (cfg.raytrans_posenc) {     // cond_nerf.py:118-127: the table is built on the host in float64, as the reference does
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 pe = __ldg(posenc_tab + s * 4 + j);
            xr[4 * j] += pe.x; xr[4 * j + 1] += pe.y; xr[4 * j + 2] += pe.z; xr[4 * j + 3] += pe.w;
          }
        }
        TRACE_TRUNK(40);
        pk2 accrg = pk(sm.p.rgb_b[0], sm.p.rgb_b[1]), accb = pk(sm.p.rgb_b[2], 0.f);
        auto colour_half = [&](const uint32_t (&r)[32], const int base) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int o2 = base + j;
            const float4 dv = *reinterpret_cast<const float4*>(&sm.dirvec[slot][cb][ray_local][o2]);
            const float hv[4] = {fmaxf(__uint_as_float(r[j]) + dv.x, 0.f), fmaxf(__uint_as_float(r[j + 1]) + dv.y, 0.f),
                                 fmaxf(__uint_as_float(r[j + 2]) + dv.z, 0.f), fmaxf(__uint_as_float(r[j + 3]) + dv.w, 0.f)};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(sm.p.rgb_w4[o2 + t]);
              const pk2 hh = pk(hv[t], hv[t]);
              accrg = pk_fma(hh, w4.x, accrg);
              accb = pk_fma(hh, w4.y, accb);
            }
          }
        };
        colour_half(rc0, 0);
        colour_half(rc1, 32);
        TRACE_TRUNK(41);
        tc::mbar_arrive(&sm.geo_empty[slot][cb]);        // this tile's dirvec / ray records may be overwritten (geometry warp, two tiles on)
        rgb[0] = 1.f / (1.f + __expf(-pk_lo(accrg)));
        rgb[1] = 1.f / (1.f + __expf(-pk_hi(accrg)));
        rgb[2] = 1.f / (1.f + __expf(-pk_lo(accb)));
      }

      // ---------------- hand the per-sample ray-transformer inputs to the ray group of this slot
      {
        const uint32_t buf = it & 1;
        TRACE_TRUNK(5);
        mbar_wait_sleep(&sm.ray_empty[slot][buf], ((it >> 1) & 1) ^ 1, 64);
        TRACE_TRUNK(6);
        float4* h = reinterpret_cast<float4*>(&sm.hand[slot][buf][row][0]);   // 80-byte rows: conflict-free 16-byte stores
        h[0] = make_float4(xr[0], xr[1], xr[2], xr[3]);
        h[1] = make_float4(xr[4], xr[5], xr[6], xr[7]);
        h[2] = make_float4(xr[8], xr[9], xr[10], xr[11]);
        h[3] = make_float4(xr[12], xr[13], xr[14], xr[15]);
        h[4] = make_float4(rgb[0], rgb[1], rgb[2], depth_cur);
        sm.hand_nv[slot][buf][row] = nvs_cur;
        tc::mbar_arrive(&sm.ray_full[slot][buf]);       // release semantics: the stores above are visible to the waiter
      }
      TRACE_TRUNK(7);
      // (no barrier here: dirvec is double-buffered by tile parity, and the barrier at the next tile's start separates this tile's
      // reads of buffer cb from the writes of window A two tiles later)
    }
  } else {
    // ================================================================== ray group: ray transformer + compositing
    reg_dec<kRegsRay>();
    const int slot = wg - 3;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int ray_local = split_ray ? 0 : row / S, s = split_ray ? slot * kTileM + row : row - ray_local * S;
    uint32_t it = 0;
    unsigned trace_n = 0;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const int64_t tile = 2 * pair + slot;
      if (tile >= n_tiles) break;
      const int64_t ray = split_ray ? pair : tile * rays_per_tile + ray_local;
      const bool valid = ray < rays.n_rays;
      const size_t n_glob = valid ? (size_t)ray * S + s : 0;
      float rgb[3], depth_t, sigma;
      {
        const uint32_t buf = it & 1;
        TRACE_RAY(0);
        mbar_wait_sleep(&sm.ray_full[slot][buf], (it >> 1) & 1, 64);
        TRACE_RAY(1);
        switch (S) {
          case 16: ray_transformer_mma<kAct, 16>(sm, slot, buf, quarter, lane, it, trace_n); break;
          case 32: ray_transformer_mma<kAct, 32>(sm, slot, buf, quarter, lane, it, trace_n); break;
          case 64: ray_transformer_mma<kAct, 64>(sm, slot, buf, quarter, lane, it, trace_n); break;
          case 128: ray_transformer_mma<kAct, 128>(sm, slot, buf, quarter, lane, it, trace_n); break;
          default: ray_transformer_mma<kAct, 256>(sm, slot, buf, quarter, lane, it, trace_n); break;
        }
        TRACE_RAY(5);
        // back to lane = row for the compositing scan
        const float4 a4 = *reinterpret_cast<const float4*>(&sm.hand[slot][buf][row][16]);
        const float n_views_seen = sm.hand_nv[slot][buf][row];
        rgb[0] = a4.x; rgb[1] = a4.y; rgb[2] = a4.z; depth_t = a4.w;
        sigma = sm.sig[slot][row];
        tc::mbar_arrive(&sm.ray_empty[slot][buf]);
        if (cfg.density_maskfill && n_views_seen < 1.f) sigma = 0.f;
        if (!valid) sigma = 0.f;
      }
      TRACE_RAY(6);
      if (aux && valid) {
        float4* a4 = reinterpret_cast<float4*>(aux) + n_glob;
        *a4 = make_float4(rgb[0], rgb[1], rgb[2], sigma);
      }

      // ---------------- alpha compositing along the ray (nerf.py:101-124, wo_render_interval)
      {
        const float part_in[5] = {rgb[0], rgb[1], rgb[2], depth_t, 1.f};
        float outv[5];
        bool writer;
        switch (S) {
          case 16: writer = composite_rows<16>(sm, slot, quarter, lane, it, sigma, valid, part_in, outv); break;
          case 32: writer = composite_rows<32>(sm, slot, quarter, lane, it, sigma, valid, part_in, outv); break;
          case 64: writer = composite_rows<64>(sm, slot, quarter, lane, it, sigma, valid, part_in, outv); break;
          case 128: writer = composite_rows<128>(sm, slot, quarter, lane, it, sigma, valid, part_in, outv); break;
          default: writer = composite_rows<256>(sm, slot, quarter, lane, it, sigma, valid, part_in, outv); break;
        }
        if (writer && valid) {
          const float bg = setbg_opaque ? 1.f - outv[4] : 0.f;
          out_rgb[ray * 3 + 0] = outv[0] + bg;
          out_rgb[ray * 3 + 1] = outv[1] + bg;
          out_rgb[ray * 3 + 2] = outv[2] + bg;
          out_depth[ray] = outv[3];
          out_opacity[ray] = outv[4];
        }
        TRACE_RAY(7);
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
  auto with_bias = [&](std::vector<double> t, int col, const float* b) {   // bias as the weight column of a constant-1 input
    for (int n = 0; n < 128; ++n) t[(size_t)n * 64 + col] = b[n];
    return t;
  };
  put_chunk(buf, 0, 128, with_bias(slice(P + off.gate_w, 128, kCond, 0, kCond), kCond, P + off.gate_b));
  put_chunk(buf, 1, 128, with_bias(slice(P + off.pts_w[0], 128, kEnc, 0, kEnc), kEnc, P + off.pts_b[0]));
  for (int l = 1; l <= 4; ++l) {
    put_chunk(buf, 2 * l, 128, slice(P + off.pts_w[l], 128, 128, 0, 64));
    put_chunk(buf, 2 * l + 1, 128, slice(P + off.pts_w[l], 128, 128, 64, 64));
  }
  put_chunk(buf, 10, 128, with_bias(slice(P + off.pts_w[5], 128, kEnc + 128, 0, kEnc), kEnc, P + off.pts_b[5]));
  {
    std::vector<double> t((size_t)128 * 64, 0.0);
    for (int j = 0; j < 4; ++j)
      for (int n = 0; n < 128; ++n) t[(size_t)n * 64 + 16 * j + 15] = P[off.pts_b[1 + j] + n];
    put_chunk(buf, 15, 128, t);
  }
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
  const float qscale = 0.5f * 1.4426950408889634f;    // 1/temperature (sqrt(d_k) = 2) and log2(e) for exp2
  for (int o = 0; o < 16; ++o)
    for (int i = 0; i < kWRow; ++i) {
      const bool in = i < 16;
      const float w3[3] = {in ? P[off.att_q + o * 16 + i] * qscale : 0.f, in ? P[off.att_k + o * 16 + i] : 0.f,
                           in ? P[off.att_v + o * 16 + i] : 0.f};
      for (int m = 0; m < 3; ++m) {
        tp.wqkv_h[16 * m + o][i] = __float2half_rn(w3[m]);
        tp.wqkv_l[16 * m + o][i] = __float2half_rn(w3[m] - __half2float(tp.wqkv_h[16 * m + o][i]));
      }
      tp.fc_h[o][i] = __float2half_rn(in ? P[off.att_fc + o * 16 + i] : 0.f);
      tp.oa0_h[o][i] = __float2half_rn(in ? P[off.oa0_w + o * 16 + i] : 0.f);
    }
  memcpy(tp.ln_w, P + off.ln_w, sizeof(tp.ln_w));
  memcpy(tp.ln_b, P + off.ln_b, sizeof(tp.ln_b));
  memcpy(tp.oa0_b, P + off.oa0_b, sizeof(tp.oa0_b));
  memcpy(tp.oa2_w, P + off.oa2_w, sizeof(tp.oa2_w));
  tp.oa2_b = P[off.oa2_b];
  memcpy(tp.alpha_b, P + off.alpha_b, sizeof(tp.alpha_b));
  for (int o = 0; o < 64; ++o) {
    tp.rgb_w4[o][0] = P[off.rgb_w + 0 * 64 + o];
    tp.rgb_w4[o][1] = P[off.rgb_w + 1 * 64 + o];
    tp.rgb_w4[o][2] = P[off.rgb_w + 2 * 64 + o];
    tp.rgb_w4[o][3] = 0.f;
  }
  memcpy(tp.rgb_b, P + off.rgb_b, sizeof(tp.rgb_b));

  DecoderWeightsTC* w = new DecoderWeightsTC();
  MNF_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&w->packed), kPackedBytes));
  MNF_CUDA_TRY(cudaMemcpy(w->packed, buf.data(), kPackedBytes, cudaMemcpyHostToDevice));
  MNF_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&w->params), sizeof(TcParams)));
  MNF_CUDA_TRY(cudaMemcpy(w->params, &tp, sizeof(TcParams), cudaMemcpyHostToDevice));
  {
    std::vector<float> tab((size_t)kMaxSamples * 16);
    for (int pos = 0; pos < kMaxSamples; ++pos)
      for (int j = 0; j < 16; ++j) {
        const double ang = (double)pos / pow(10000.0, 2.0 * (double)(j / 2) / 16.0);
        tab[(size_t)pos * 16 + j] = (float)((j & 1) ? cos(ang) : sin(ang));
      }
    MNF_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&w->posenc), tab.size() * sizeof(float)));
    MNF_CUDA_TRY(cudaMemcpy(w->posenc, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  *out = w;
  return MNF_OK;
}

void decoder_tc_free(DecoderWeightsTC* w) {
  if (!w) return;
  cudaFree(w->packed);
  cudaFree(w->params);
  cudaFree(w->posenc);
  delete w;
}

bool decoder_tc_supports(const mnf_decoder_cfg& cfg) {
  const int S = cfg.n_samples;
  return S == 16 || S == 32 || S == 64 || S == 128 || S == 256;
}

int launch_decoder_tc(const DevCams& cams, const DevRays& rays, const mnf_decoder_cfg& cfg, const DecoderWeightsTC* w,
                      const HeadParams*, const __half* cond_f16, int setbg_opaque, float* out_rgb, float* out_depth,
                      float* out_opacity, float* aux, cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  if (!w) { set_error("tcgen05 decoder weights not packed"); return MNF_ESTATE; }
  static PerDevice<int> n_sm_dev;
  int& n_sm = n_sm_dev.cur();
  const size_t smem = sizeof(TcSmem) + 1024;
  if (n_sm == 0) {
    int dev = 0;
    MNF_CUDA_TRY(cudaGetDevice(&dev));
    MNF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    MNF_CUDA_TRY(cudaFuncSetAttribute(decoder_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA_TRY(cudaFuncSetAttribute(decoder_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const bool split_ray = cfg.n_samples > kTileM;
  const int rays_per_tile = split_ray ? 1 : kTileM / cfg.n_samples;
  const int64_t n_tiles = split_ray ? 2 * rays.n_rays : (rays.n_rays + rays_per_tile - 1) / rays_per_tile;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const unsigned grid = (unsigned)(n_pairs < n_sm ? n_pairs : n_sm);
  if (cfg.raytrans_act == 0)
    decoder_tc_kernel<0><<<grid, kThreads, smem, s>>>(cams, rays, cfg, w->packed, w->params, w->posenc, cond_f16, setbg_opaque, out_rgb,
                                                      out_depth, out_opacity, aux);
  else
    decoder_tc_kernel<1><<<grid, kThreads, smem, s>>>(cams, rays, cfg, w->packed, w->params, w->posenc, cond_f16, setbg_opaque, out_rgb,
                                                      out_depth, out_opacity, aux);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf

// Debug aid (not part of the reference-facing ABI): arm / disarm the decoder timeline trace.  buf = device array of
// `cap` u64 records ([clock:40 | role:4 | slot:4 | event:8 | iteration:8]); buf = NULL disarms.  Synchronises.
extern "C" int32_t mnf_debug_decoder_trace(void* buf, int32_t cap) {
  using namespace mnf;
  unsigned long long* p = reinterpret_cast<unsigned long long*>(buf);
  unsigned int c = buf ? (unsigned)cap : 0u;
#ifndef MNF_DECODER_TRACE
  if (buf) { set_error("library built without -DMNF_DECODER_TRACE"); return MNF_EUNSUPPORTED; }
#endif
  MNF_CUDA_TRY(cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)));
  MNF_CUDA_TRY(cudaMemcpyToSymbol(g_trace_cap, &c, sizeof(c)));
  MNF_CUDA_TRY(cudaDeviceSynchronize());
  return MNF_OK;
}
