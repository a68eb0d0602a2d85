// K-gather v6: the bilinear blend of the epipolar feature gather on the 5th-gen tensor cores, fed by TMA.
//
// Replaces MatchNeRF.query_cond_info (models/matchnerf.py:209-293) + the ray casting / depth sampling / projection it depends
// on (misc/camera.py:255-286, :351-379; matchnerf.py:163-181) for CONTIGUOUS ray ranges (render_by_slices / full images).
//
// Why: v3 (gather.cu) sits at ~57 % of the CUDA-core FMA roof -- 6144 blend + 2304 product multiply-adds per sample on the
// half-rate fp16 pipe -- so no CUDA-core formulation can be 2x faster.  The blend is a GEMM if the samples of a tile share
// their texels:   blended[sample][channel] = sum_k W[sample][k] * T[k][channel]
// with T = the K texels of the tile's footprint in a source feature map and W = 4 bilinear weights per row.  A tile is a 16 x 8
// PIXEL block (128 rays) at ONE depth sample: adjacent pixels project 1/4 (1/8) texel apart at the fine (coarse) scale, so the
// footprint of the whole block is a box of <= 8 x 4 (4 x 4) texels = K 32 (16):
//   * the box [BY][BX][256 channels] arrives by ONE 4-D TMA tensor-map copy per (view, scale) (cp.async.bulk.tensor, SWIZZLE_128B),
//     laid out by the TMA unit as the MN-major B operand [channel block][texel][64 channels];
//   * the 128 geometry threads (thread = ray) write their rows of W (fp16, 4 non-zeros) as the K-major A operand;
//   * tcgen05.mma (M 128 samples, N 64 channels, K 16 per step) accumulates the blended features of the 3 views in fp32 in
//     tensor memory, 64 channel positions at a time, double buffered;
//   * the 128 product threads (thread = sample = TMEM lane) read their row with tcgen05.ld and form the 9 pair products per channel
//     group in registers with packed fp32 FMAs -- no cross-lane reduction at all -- then the cosines, and write the 64-byte
//     conditioning row together with the colours / masks the geometry threads handed over.
// Per sample: 9 tensor-pipe cycles, ~36 packed-FMA + ~12 tcgen05.ld warp-instructions instead of v3's 367 warp-instructions.
//
// A tile whose footprint does not fit the boxes at some depth (strong foreshortening, points behind a camera) is put on a list
// and recomputed by the v3 kernel afterwards (gather.cu, tile-list mode); explicit ray lists / sample points always use v3.
#include <cuda.h>

#include <cstdlib>

#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {

constexpr int kTW = 16, kTH = 8;                       // pixel tile: 16 x 8 = 128 rays
constexpr int kRays = kTW * kTH;
constexpr int kBX0 = 4, kBY0 = 4, kK0 = kBX0 * kBY0;   // coarse-scale box: 16 texels
constexpr int kBX1 = 8, kBY1 = 4, kK1 = kBX1 * kBY1;   // fine-scale box:   32 texels
constexpr int kTexB = kFeatCh * 2;                     // 512 B per texel
constexpr int kT0 = kK0 * kTexB, kT1 = kK1 * kTexB;    // bytes of one view's box
constexpr int kTSlot = kViews * (kT0 + kT1);           // all boxes of one depth sample: 73,728 B
constexpr int kWTile = kRays * 128;                    // one [128 rows][64 fp16] SWIZZLE_128B tile: 16 KB
constexpr int kChunk = 64;                             // channel positions per accumulator chunk
constexpr int kDCols = kViews * kChunk;                // 192 TMEM columns per accumulator buffer
constexpr int kThreadsTc = 416;                        // warps 0-3 geometry, 4-7 / 8-11 products (even / odd accumulator chunks), 12 MMA issue
constexpr int kMmaWarp = 12;
constexpr int kColX = 2 * kDCols;                      // 32 spare TMEM columns: coarse-group partial sums handed from product set A to set B

struct GtSmem {
  alignas(1024) unsigned char T[2][kTSlot];            // [depth slot]: coarse v0 v1 v2, then fine v0 v1 v2
  alignas(1024) unsigned char Wt[2][2][kWTile];        // [unit slot][tile]; coarse: tile 0 holds views 0,1,2 at k 0,16,32;
                                                       //                     fine: tile 0 views 0,1 at k 0,32, tile 1 view 2 at k 0
  alignas(16) float hand[2][kRays][12];                // geometry -> product threads: colours (9) + masks (3) of a sample
  alignas(8) float sims[kRays][10];                    // product threads: the 10 cosines of the depth sample in flight (thread-private rows)
  int red[4][24];                                      // per geometry warp: box extents (min x, max x, min y, max y) x 6
  alignas(8) uint64_t t_full[2], t_free[2], w_ready[2], w_free[2], d_full[2], d_free[2], h_full[2], h_free[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void geo_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          tc::smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(tc::smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float2 ffma2_(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// orders uses of a tcgen05.ld destination after the (single) tcgen05.wait::ld that precedes this statement; emits no instruction
__device__ __forceinline__ void tie16(uint32_t (&r)[16]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                    "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
// mean over the three pairs of <A,B> / (max(|A|,eps) max(|B|,eps))   (models/matchnerf.py:268-271); same form as gather.cu
__device__ __forceinline__ float mean_cosine9(const float2 (&q)[9]) {
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const float ab = q[3 * p].x + q[3 * p].y, aa = q[3 * p + 1].x + q[3 * p + 1].y, bb = q[3 * p + 2].x + q[3 * p + 2].y;
    acc += ab * rsqrt_approx(fmaxf(aa, 1e-16f)) * rsqrt_approx(fmaxf(bb, 1e-16f));   // operands >= 1e-16: no denormal handling needed
  }
  return acc * (1.0f / 3.0f);
}

// One run of 8 packed positions = channels 4l..4l+3 of half 0 then of half 1 (pack.cu, packing v3) of the three views:
// pairs (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1), each as (dot, |a|^2, |b|^2), accumulated as packed pairs.
__device__ __forceinline__ void run_products(const uint32_t* a, const uint32_t* b, const uint32_t* c, float2 (&q)[9]) {
  const float2 a0 = make_float2(__uint_as_float(a[0]), __uint_as_float(a[1])), a1 = make_float2(__uint_as_float(a[2]), __uint_as_float(a[3]));
  const float2 a2 = make_float2(__uint_as_float(a[4]), __uint_as_float(a[5])), a3 = make_float2(__uint_as_float(a[6]), __uint_as_float(a[7]));
  const float2 b0 = make_float2(__uint_as_float(b[0]), __uint_as_float(b[1])), b1 = make_float2(__uint_as_float(b[2]), __uint_as_float(b[3]));
  const float2 b2 = make_float2(__uint_as_float(b[4]), __uint_as_float(b[5])), b3 = make_float2(__uint_as_float(b[6]), __uint_as_float(b[7]));
  const float2 c0 = make_float2(__uint_as_float(c[0]), __uint_as_float(c[1])), c1 = make_float2(__uint_as_float(c[2]), __uint_as_float(c[3]));
  const float2 c2 = make_float2(__uint_as_float(c[4]), __uint_as_float(c[5])), c3 = make_float2(__uint_as_float(c[6]), __uint_as_float(c[7]));
  q[0] = ffma2_(a1, b1, ffma2_(a0, b0, q[0])); q[1] = ffma2_(a1, a1, ffma2_(a0, a0, q[1])); q[2] = ffma2_(b1, b1, ffma2_(b0, b0, q[2]));
  q[3] = ffma2_(a3, c1, ffma2_(a2, c0, q[3])); q[4] = ffma2_(a3, a3, ffma2_(a2, a2, q[4])); q[5] = ffma2_(c1, c1, ffma2_(c0, c0, q[5]));
  q[6] = ffma2_(b3, c3, ffma2_(b2, c2, q[6])); q[7] = ffma2_(b3, b3, ffma2_(b2, b2, q[7])); q[8] = ffma2_(c3, c3, ffma2_(c2, c2, q[8]));
}

struct TapTc {          // one (view, scale) of one sample
  int x0, y0;           // texel of tap 00
  uint32_t w01, w23;    // fp16 weights (w00, w01), (w10, w11)
  int dx, dy;           // 1 when the +1 tap carries weight
};

__device__ __forceinline__ TapTc make_tap_tc(float gx, float gy, int w, int h) {
  const float ix = grid_unnormalize(gx, w), iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = ix - x0f, fy = iy - y0f;
  const __half2 a = __floats2half2_rn((1.f - fx) * (1.f - fy), fx * (1.f - fy));
  const __half2 b = __floats2half2_rn((1.f - fx) * fy, fx * fy);
  TapTc t;
  t.x0 = (int)x0f; t.y0 = (int)y0f;
  t.w01 = *reinterpret_cast<const uint32_t*>(&a); t.w23 = *reinterpret_cast<const uint32_t*>(&b);
  t.dx = fx > 0.f ? 1 : 0; t.dy = fy > 0.f ? 1 : 0;
  return t;
}

__device__ __forceinline__ void bilinear_setup_tc(float gx, float gy, int w, int h, uint32_t& off, float& fx, float& fy) {
  const float ix = grid_unnormalize(gx, w), iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  fx = ix - x0f; fy = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const uint32_t dx = x0 + 1 <= w - 1 ? 1u : 0u, dy = y0 + 1 <= h - 1 ? 1u : 0u;
  off = (uint32_t)(y0 * w + x0) | (dx << 30) | (dy << 31);
}

// zero `n16` 16-byte chunks starting at chunk `c0` of row `row` of a SWIZZLE_128B [128][64 fp16] tile
__device__ __forceinline__ void zero_row_chunks(unsigned char* tile, int row, int c0, int n16) {
  unsigned char* base = tile + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < n16) *reinterpret_cast<uint4*>(base + ((((c0 + c) ^ (row & 7)) & 7) << 4)) = make_uint4(0u, 0u, 0u, 0u);
}
// row_base = tile + (row >> 3) * 1024 + (row & 7) * 128, r7 = row & 7: one predicated 2-byte store, no branch
__device__ __forceinline__ void put_weight(unsigned char* row_base, int r7, int k, unsigned short w, bool on) {
  unsigned char* p = row_base + ((((k >> 3) ^ r7) & 7) << 4) + (k & 7) * 2;
  if (on) *reinterpret_cast<unsigned short*>(p) = w;
}

}  // namespace

static_assert(sizeof(GtSmem) + 1024 <= 232448, "gather_tc shared memory exceeds the 227 KB a CTA can have");

struct GatherTcArgs {
  int S, h0, w0, h1, w1;
  int tiles_x, band0, n_tiles;           // pixel-tile grid covering the ray range: tile = band * tiles_x + tx
  int* overflow_list;                    // [n_tiles] tile ids whose footprint left the boxes at some depth
  int* overflow_count;
};

__global__ void __launch_bounds__(kThreadsTc, 1)
gather_tc_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const GatherTcArgs args,
                 const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  extern __shared__ unsigned char smem_dyn[];
  GtSmem& sm = *reinterpret_cast<GtSmem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = args.S;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&sm.t_full[i], 1);
      tc::mbar_init(&sm.t_free[i], 1);
      tc::mbar_init(&sm.w_ready[i], kRays);
      tc::mbar_init(&sm.w_free[i], 1);
      tc::mbar_init(&sm.d_full[i], 1);
      tc::mbar_init(&sm.d_free[i], kRays);
      tc::mbar_init(&sm.h_full[i], kRays);
      tc::mbar_init(&sm.h_free[i], kRays);
    }
    tc::fence_mbar_init();
  }
  if (warp == kMmaWarp) tc::tmem_alloc<512>(&sm.tmem_base);
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  // counters shared by construction: every role walks the same (tile, depth sample, scale) sequence
  uint32_t gs = 0;                       // depth samples processed by this CTA so far (selects T / hand slots)

  if (warp < 4) {
    // ============================================================================================ geometry threads (thread = ray)
    const int px = tid & (kTW - 1), py = tid >> 4;
    const int HW = cams.H * cams.W;
    for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
      const int band = args.band0 + tile / args.tiles_x, tx = tile - (tile / args.tiles_x) * args.tiles_x;
      const int x = tx * kTW + px, y = band * kTH + py;
      const bool inside = x < cams.W && y < cams.H;
      const int64_t pix = (int64_t)min(y, cams.H - 1) * cams.W + min(x, cams.W - 1);
      const int64_t rel = pix - rays.first_ray;
      const bool valid = inside && rel >= 0 && rel < rays.n_rays;
      float o[3], d[3];
      cast_ray(cams, pix, o, d);
      bool overflow = false;
      for (int s = 0; s < S; ++s, ++gs) {
        const uint32_t slot = gs & 1u, par = (gs >> 1) & 1u;
        // ---------------- per-sample geometry (identical arithmetic to gather.cu: the mask decisions hang on it)
        const float u = (rays.jitter && valid) ? rays.jitter[rel * S + s] : 0.f;
        const float t = sample_depth(cams, s, S, u);
        float p[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));
        TapTc tap[kViews][2];
        float colmask[12];
#pragma unroll
        for (int v = 0; v < kViews; ++v) {
          float uu, vv, zz;
          project_ndc(cams, v, p, uu, vv, zz);
          const float gx = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f), gy = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
          colmask[9 + v] = (gx > -1.0f && gx < 1.0f && gy > -1.0f && gy < 1.0f) ? 1.f : 0.f;
          tap[v][0] = make_tap_tc(gx, gy, args.w0, args.h0);
          tap[v][1] = make_tap_tc(gx, gy, args.w1, args.h1);
          uint32_t coff;
          float fx, fy;
          bilinear_setup_tc(gx, gy, cams.W, cams.H, coff, fx, fy);
          const uint32_t o00 = coff & 0x3fffffffu, dx = (coff >> 30) & 1u, dy = coff >> 31;
          const float4* pr = images + (size_t)v * HW + o00;
          const float4 c00 = __ldg(pr), c01 = __ldg(pr + dx), c10 = __ldg(pr + dy * cams.W), c11 = __ldg(pr + dy * cams.W + dx);
          const float wa0 = (1.f - fx) * (1.f - fy), wb0 = fx * (1.f - fy), wa1 = (1.f - fx) * fy, wb1 = fx * fy;
          colmask[3 * v + 0] = (c00.x * wa0 + c01.x * wb0) + (c10.x * wa1 + c11.x * wb1);
          colmask[3 * v + 1] = (c00.y * wa0 + c01.y * wb0) + (c10.y * wa1 + c11.y * wb1);
          colmask[3 * v + 2] = (c00.z * wa0 + c01.z * wb0) + (c10.z * wa1 + c11.z * wb1);
        }
        // ---------------- footprint boxes of the tile: min / max texel over the 128 rays, per (view, scale)
        int bx[kViews][2], by[kViews][2];
        {
          int e[24];
#pragma unroll
          for (int v = 0; v < kViews; ++v)
#pragma unroll
            for (int sc = 0; sc < 2; ++sc) {
              const int k = (v * 2 + sc) * 4;
              e[k + 0] = __reduce_min_sync(0xffffffffu, tap[v][sc].x0);
              e[k + 1] = __reduce_max_sync(0xffffffffu, tap[v][sc].x0 + tap[v][sc].dx);
              e[k + 2] = __reduce_min_sync(0xffffffffu, tap[v][sc].y0);
              e[k + 3] = __reduce_max_sync(0xffffffffu, tap[v][sc].y0 + tap[v][sc].dy);
            }
          geo_barrier();                                 // the previous sample's extents have been consumed by every warp
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 24; k += 4) *reinterpret_cast<int4*>(&sm.red[warp][k]) = make_int4(e[k], e[k + 1], e[k + 2], e[k + 3]);
          }
          geo_barrier();
#pragma unroll
          for (int v = 0; v < kViews; ++v)
#pragma unroll
            for (int sc = 0; sc < 2; ++sc) {
              const int k = (v * 2 + sc) * 4;
              int4 m = *reinterpret_cast<const int4*>(&sm.red[0][k]);
#pragma unroll
              for (int w2 = 1; w2 < 4; ++w2) {
                const int4 n = *reinterpret_cast<const int4*>(&sm.red[w2][k]);
                m.x = min(m.x, n.x); m.y = max(m.y, n.y); m.z = min(m.z, n.z); m.w = max(m.w, n.w);
              }
              bx[v][sc] = m.x; by[v][sc] = m.z;
              overflow |= (m.y - m.x >= (sc ? kBX1 : kBX0)) || (m.w - m.z >= (sc ? kBY1 : kBY0));
            }
        }
        // ---------------- TMA: the six boxes of this depth sample (one thread), into T slot `slot`
        if (tid == 0) {
          tc::mbar_wait_sleep(&sm.t_free[slot], par ^ 1u, 20);
          tc::mbar_arrive_expect_tx(&sm.t_full[slot], (uint32_t)kTSlot);
#pragma unroll
          for (int v = 0; v < kViews; ++v) {
            tma_load_4d(sm.T[slot] + v * kT0, &map0, 0, bx[v][0], v * args.h0 + by[v][0], 0, &sm.t_full[slot]);
            tma_load_4d(sm.T[slot] + kViews * kT0 + v * kT1, &map1, 0, bx[v][1], v * args.h1 + by[v][1], 0, &sm.t_full[slot]);
          }
        }
        // ---------------- W rows of both scales (K-major A operand), unit slots alternate coarse / fine
#pragma unroll
        for (int sc = 0; sc < 2; ++sc) {
          const uint32_t gu = gs * 2u + (uint32_t)sc, us = gu & 1u, upar = (gu >> 1) & 1u;
          tc::mbar_wait_sleep(&sm.w_free[us], upar ^ 1u, 20);
          const int K = sc ? kK1 : kK0, BX = sc ? kBX1 : kBX0, BY = sc ? kBY1 : kBY0;
#pragma unroll
          for (int v = 0; v < kViews; ++v) {
            unsigned char* tile_w = sm.Wt[us][(sc == 1 && v == 2) ? 1 : 0];
            const int kbase = (sc == 1 && v == 2) ? 0 : v * K;
            zero_row_chunks(tile_w, tid, kbase >> 3, K >> 3);
            const TapTc& tp = tap[v][sc];
            const int lx = tp.x0 - bx[v][sc], ly = tp.y0 - by[v][sc];
            const unsigned short w00 = (unsigned short)(tp.w01 & 0xffffu), w01 = (unsigned short)(tp.w01 >> 16);
            const unsigned short w10 = (unsigned short)(tp.w23 & 0xffffu), w11 = (unsigned short)(tp.w23 >> 16);
            const bool in0 = lx < BX && ly < BY, inx = lx + 1 < BX, iny = ly + 1 < BY;   // (lx, ly >= 0 by construction of the box)
            unsigned char* row_base = tile_w + (tid >> 3) * 1024 + (tid & 7) * 128;
            const int k00 = kbase + ly * BX + lx;
            put_weight(row_base, tid & 7, k00, w00, in0);
            put_weight(row_base, tid & 7, k00 + 1, w01, tp.dx && inx && ly < BY);
            put_weight(row_base, tid & 7, k00 + BX, w10, tp.dy && iny && lx < BX);
            put_weight(row_base, tid & 7, k00 + BX + 1, w11, tp.dx && tp.dy && inx && iny);
          }
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&sm.w_ready[us]);
        }
        // ---------------- colours / masks to the product threads
        tc::mbar_wait_sleep(&sm.h_free[slot], par ^ 1u, 20);
        {
          float4* h = reinterpret_cast<float4*>(&sm.hand[slot][tid][0]);
          h[0] = make_float4(colmask[0], colmask[1], colmask[2], colmask[3]);
          h[1] = make_float4(colmask[4], colmask[5], colmask[6], colmask[7]);
          h[2] = make_float4(colmask[8], colmask[9], colmask[10], colmask[11]);
        }
        tc::mbar_arrive(&sm.h_full[slot]);
      }
      if (tid == 0 && overflow) args.overflow_list[atomicAdd(args.overflow_count, 1)] = tile;   // every thread holds the same flag
    }
  } else if (warp < kMmaWarp) {
    // ============================================================================================ product threads (thread = sample)
    // ONE compact loop over accumulator chunks (64 packed positions x 3 views), four steps of 16 positions each: ~250 instructions
    // that stay in the instruction cache (the first version unrolled all 32 steps of a depth sample: 36 KB of straight-line code,
    // and ncu showed `no instruction` as the top stall of these warps).  The tcgen05.ld of step t + 1 -- also across chunk, scale,
    // sample and tile boundaries -- is in flight while step t is multiplied out.
    // TWO sets of 128 product threads share the work by accumulator buffer: set A (warps 4-7) multiplies out the even chunks
    // (buffer 0), set B (warps 8-11) the odd ones (buffer 1) -- a single warp per SM sub-partition ran this loop at ~0.19 IPC (ncu:
    // fixed-latency `wait` stalls between dependent FFMA2s and tcgen05.ld round trips), so a second warp per sub-partition nearly
    // doubles the rate.  A fine cosine group (32 positions) lies inside a chunk; a coarse group (128 positions) spans an even and
    // an odd chunk: set A hands its nine partial sums to set B through 16 spare tensor-memory columns (same lane = same sample)
    // and a 64-thread named barrier per lane quarter.  Set B assembles and writes the conditioning row.
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const uint32_t set = (uint32_t)(warp >> 2) - 1u;                  // 0: even chunks, 1: odd chunks
    const uint32_t tb = tmem + ((uint32_t)(quarter * 32) << 16);
    const int px = row & (kTW - 1), py = row >> 4;
    float* const simrow = &sm.sims[row][0];          // the 10 cosines of the depth sample in flight (one row per sample, written by both sets)
    const int my_tiles = args.n_tiles > (int)blockIdx.x ? (args.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t n_chunks = (uint32_t)my_tiles * (uint32_t)S * 8u;
    const uint32_t d0 = tb + set * kDCols;           // this set's accumulator buffer
    auto pair_arrive = [&](int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); };
    auto pair_sync = [&](int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); };
    uint32_t ra[2][16], rb[2][16], rc[2][16];
    float2 q[9];
    if (n_chunks > 0) {
      tc::mbar_wait(&sm.d_full[set], 0u);
      tc::tc_fence_after_sync();
      tc::tmem_ld16(d0, ra[0]);
      tc::tmem_ld16(d0 + kChunk, rb[0]);
      tc::tmem_ld16(d0 + 2 * kChunk, rc[0]);
      tc::tmem_wait_ld(ra[0]); tie16(rb[0]); tie16(rc[0]);
    }
    int tile = blockIdx.x, s = 0;
#pragma unroll 1
    for (uint32_t g = set; g < n_chunks; g += 2u) {
      const int sc = (int)((g >> 2) & 1u), c = (int)(g & 3u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cur = j & 1, nxt = cur ^ 1;
        if (j != 3) {                                 // next step lies in the same chunk
          tc::tmem_ld16(d0 + 16 * (j + 1), ra[nxt]);
          tc::tmem_ld16(d0 + kChunk + 16 * (j + 1), rb[nxt]);
          tc::tmem_ld16(d0 + 2 * kChunk + 16 * (j + 1), rc[nxt]);
        } else {
          // every load of this chunk completed a step ago: hand the buffer back, then look at this set's next chunk
          tc::tc_fence_before_sync();
          tc::mbar_arrive(&sm.d_free[set]);
          if (g + 2u < n_chunks) {
            const uint32_t g2 = g + 2u;
            tc::mbar_wait_sleep(&sm.d_full[set], (g2 >> 1) & 1u, 20);
            tc::tc_fence_after_sync();
            tc::tmem_ld16(d0, ra[nxt]);
            tc::tmem_ld16(d0 + kChunk, rb[nxt]);
            tc::tmem_ld16(d0 + 2 * kChunk, rc[nxt]);
          }
        }
        if (j % 2 == 0 && (sc == 1 || j == 0)) {      // a new cosine (partial) group starts: fine = 32 positions, coarse = this chunk's 64
#pragma unroll
          for (int i = 0; i < 9; ++i) q[i] = make_float2(0.f, 0.f);
        }
        run_products(ra[cur], rb[cur], rc[cur], q);
        run_products(ra[cur] + 8, rb[cur] + 8, rc[cur] + 8, q);
        if (j % 2 == 1) {
          if (sc == 1) {
            simrow[2 + c * 2 + j / 2] = mean_cosine9(q);
          } else if (j == 3) {
            const uint32_t xcol = tb + kColX + (uint32_t)(c >> 1) * 16u;
            const int bar_id = 2 + quarter * 2 + (c >> 1);
            if (set == 0u) {                          // first half of a coarse group: nine sums -> tensor memory -> set B
              uint32_t x[16];
#pragma unroll
              for (int i = 0; i < 9; ++i) x[i] = __float_as_uint(q[i].x + q[i].y);
#pragma unroll
              for (int i = 9; i < 16; ++i) x[i] = 0u;
              tc::tmem_wait_ld(ra[nxt]); tie16(rb[nxt]); tie16(rc[nxt]);    // (the prefetch shares the wait below with nothing else)
              tc::tmem_st16(xcol, x);
              tc::tmem_wait_st();
              tc::tc_fence_before_sync();
              pair_arrive(bar_id);
            } else {                                  // second half: add set A's sums, then the cosine
              pair_sync(bar_id);
              tc::tc_fence_after_sync();
              uint32_t x[16];
              tc::tmem_ld16(xcol, x);
              tc::tmem_wait_ld(x); tie16(ra[nxt]); tie16(rb[nxt]); tie16(rc[nxt]);
#pragma unroll
              for (int i = 0; i < 9; ++i) q[i] = make_float2(q[i].x + q[i].y, __uint_as_float(x[i]));
              simrow[c >> 1] = mean_cosine9(q);
            }
          }
        }
        tc::tmem_wait_ld(ra[nxt]); tie16(rb[nxt]); tie16(rc[nxt]);
      }
      if (sc == 1 && c == 2) pair_arrive(10 + quarter);        // set A: its fine cosines of this depth sample are in simrow
      if (sc == 1 && c == 3) {
        // ---------------- assemble the conditioning row: feat_info[10], color_info[9], mask_info[3] (cond_nerf.py:59)   (set B)
        pair_sync(10 + quarter);
        const uint32_t gsl = g >> 3, slot = gsl & 1u, par = (gsl >> 1) & 1u;
        const int band = args.band0 + tile / args.tiles_x, tx = tile - (tile / args.tiles_x) * args.tiles_x;
        const int x = tx * kTW + px, y = band * kTH + py;
        const int64_t rel = (int64_t)y * cams.W + x - rays.first_ray;
        const bool valid = x < cams.W && y < cams.H && rel >= 0 && rel < rays.n_rays;
        tc::mbar_wait_sleep(&sm.h_full[slot], par, 20);
        const float4* h = reinterpret_cast<const float4*>(&sm.hand[slot][row][0]);
        const float4 h0 = h[0], h1 = h[1], h2 = h[2];
        tc::mbar_arrive(&sm.h_free[slot]);
        if (valid) {
          const float2 s0 = *reinterpret_cast<const float2*>(simrow), s1 = *reinterpret_cast<const float2*>(simrow + 2);
          const float2 s2 = *reinterpret_cast<const float2*>(simrow + 4), s3 = *reinterpret_cast<const float2*>(simrow + 6);
          const float2 s4 = *reinterpret_cast<const float2*>(simrow + 8);
          const float vals[24] = {s0.x, s0.y, s1.x, s1.y, s2.x, s2.y, s3.x, s3.y, s4.x, s4.y,
                                  h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w, h2.x, h2.y, h2.z, h2.w, 0.f, 0.f};
          const size_t n = (size_t)rel * S + s;
          if (cond_f16) {
            uint4 o4[4];
            uint32_t* o = reinterpret_cast<uint32_t*>(o4);
#pragma unroll
            for (int i = 0; i < 12; ++i) {
              const __half2 hh = __floats2half2_rn(vals[2 * i], vals[2 * i + 1]);
              o[i] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            o[12] = o[13] = o[14] = o[15] = 0u;
            uint4* dst = reinterpret_cast<uint4*>(cond_f16 + n * kCondPad);
            dst[0] = o4[0]; dst[1] = o4[1]; dst[2] = o4[2]; dst[3] = o4[3];
          }
          if (cond_f32) {
            float2* dst = reinterpret_cast<float2*>(cond_f32 + n * kCond);      // 88-byte rows: 8-byte aligned
#pragma unroll
            for (int i = 0; i < 11; ++i) dst[i] = make_float2(vals[2 * i], vals[2 * i + 1]);
          }
        }
        if (++s == S) { s = 0; tile += gridDim.x; }
      }
    }
  } else {
    // ============================================================================================ MMA issuer (one elected lane)
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::umma_idesc_f16(128, kChunk, 1u);          // B MN-major: [texel][64 channels]
    for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
      for (int s = 0; s < S; ++s, ++gs) {
        const uint32_t slot = gs & 1u, par = (gs >> 1) & 1u;
        tc::mbar_wait_sleep(&sm.t_full[slot], par, 20);
#pragma unroll
        for (int sc = 0; sc < 2; ++sc) {
          const uint32_t gu = gs * 2u + (uint32_t)sc, us = gu & 1u, upar = (gu >> 1) & 1u;
          tc::mbar_wait_sleep(&sm.w_ready[us], upar, 20);
          tc::tc_fence_after_sync();
          const int K = sc ? kK1 : kK0;
          const uint32_t tview = sc ? kT1 : kT0;
          const uint32_t tbase = tc::smem_u32(sm.T[slot]) + (sc ? kViews * kT0 : 0);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t g = gu * 4u + (uint32_t)c, db = g & 1u, dpar = (g >> 1) & 1u;
            tc::mbar_wait_sleep(&sm.d_free[db], dpar ^ 1u, 20);
            tc::tc_fence_after_sync();
            if (leader) {
#pragma unroll
              for (int v = 0; v < kViews; ++v) {
                const uint32_t wt = tc::smem_u32(sm.Wt[us][(sc == 1 && v == 2) ? 1 : 0]);
                const int kbase = (sc == 1 && v == 2) ? 0 : v * K;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  if (ks * 16 < K) {
                    const uint64_t ad = tc::umma_desc_sw128(wt + (uint32_t)(kbase + ks * 16) * 2u);
                    const uint64_t bd = tc::umma_desc_sw128_mn(tbase + v * tview + (uint32_t)c * (uint32_t)(K * 128) + (uint32_t)ks * 2048u,
                                                               (uint32_t)(K * 128));
                    tc::umma_ss(tmem + db * kDCols + v * kChunk, ad, bd, idesc, ks > 0 ? 1u : 0u);
                  }
                }
              }
              tc::umma_commit(&sm.d_full[db]);
            }
            __syncwarp();
          }
          if (leader) {
            tc::umma_commit(&sm.w_free[us]);
            if (sc == 1) tc::umma_commit(&sm.t_free[slot]);
          }
          __syncwarp();
        }
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------ host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 4-D view of a packed feature map [V][h][w][256 fp16]: (64 positions of a 128-byte block, x, V*h rows, 4 blocks); a box of
// (64, BX, BY, 4) lands in shared memory as [block][BY][BX][128 B] = the MN-major UMMA B operand [block][texel][64 positions].
int make_map(CUtensorMap* m, const __half* base, int V, int h, int w, int BX, int BY) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return MNF_ECUDA; }
  const cuuint64_t dims[4] = {64, (cuuint64_t)w, (cuuint64_t)V * h, 4};
  const cuuint64_t strides[3] = {(cuuint64_t)kTexB, (cuuint64_t)w * kTexB, 128};
  const cuuint32_t box[4] = {64, (cuuint32_t)BX, (cuuint32_t)BY, 4};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (map %dx%d, box %dx%d)", (int)r, h, w, BY, BX); return MNF_ECUDA; }
  return MNF_OK;
}

}  // namespace

int launch_gather_tc(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1, int w1,
                     const float* images, float* cond_f32, __half* cond_f16, int* scratch, int scratch_ints, cudaStream_t s,
                     DevRays* fix, int64_t* fix_blocks) {
  GatherTcArgs a{};
  a.S = S; a.h0 = h0; a.w0 = w0; a.h1 = h1; a.w1 = w1;
  a.tiles_x = (cams.W + kTW - 1) / kTW;
  const int row_first = (int)(rays.first_ray / cams.W), row_last = (int)((rays.first_ray + rays.n_rays - 1) / cams.W);
  a.band0 = row_first / kTH;
  const int n_bands = row_last / kTH - a.band0 + 1;
  a.n_tiles = a.tiles_x * n_bands;
  if (a.n_tiles + 1 > scratch_ints || (((uintptr_t)f0 | (uintptr_t)f1) & 15) != 0) return 1;     // not applicable: the caller uses v3
  a.overflow_count = scratch;
  a.overflow_list = scratch + 1;
  CUtensorMap m0, m1;
  int rc;
  if ((rc = make_map(&m0, f0, kViews, h0, w0, kBX0, kBY0))) return rc;
  if ((rc = make_map(&m1, f1, kViews, h1, w1, kBX1, kBY1))) return rc;
  static PerDevice<int> n_sm_dev;
  int& n_sm = n_sm_dev.cur();
  const size_t smem = sizeof(GtSmem) + 1024;
  if (n_sm == 0) {
    int dev = 0;
    MNF_CUDA_TRY(cudaGetDevice(&dev));
    MNF_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    MNF_CUDA_TRY(cudaFuncSetAttribute(gather_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  MNF_CUDA_TRY(cudaMemsetAsync(scratch, 0, sizeof(int), s));
  const unsigned grid = (unsigned)(a.n_tiles < n_sm ? a.n_tiles : n_sm);
  gather_tc_kernel<<<grid, kThreadsTc, smem, s>>>(cams, rays, a, m0, m1, reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  *fix = rays;
  fix->tile_list = a.overflow_list;
  fix->tile_count = a.overflow_count;
  fix->tiles_x = a.tiles_x;
  fix->band0 = a.band0;
  *fix_blocks = (int64_t)a.n_tiles * 8;
  return 0;
}

}  // namespace mnf
