// K-gather: epipolar projection + bilinear feature gather + grouped cosine similarity.
//
// Replaces MatchNeRF.query_cond_info (models/matchnerf.py:209-293) together with the ray casting /
// depth sampling / projection it depends on (misc/camera.py:255-286, :351-379; matchnerf.py:163-181).
//
// History (each step from an ncu capture, profiles/r01_ncu_summary.md):
//   v1  one sample per warp, geometry replicated in 32 lanes: 1442 warp-instructions per sample, issue-bound;
//   v2  a warp owns a quad of 4 adjacent rays, lane group = ray, geometry once per sample with lane = (ray, sample):
//       ~350 instructions per sample, 2.81 ms per 81,920 rays x 64, L1/TEX 88 % (every ray pulled its own copy of a cell);
//   v3  (this file) the lanes own the 32 slots of a texel and the warp re-uses a fetched cell for the rays of its quad.
#include <cstdlib>

#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kQuad = 4;        // rays per warp
constexpr int kChunk = 8;       // samples per geometry phase (kQuad * kChunk == 32 lanes)

// colour taps: texel index of tap 00 | dx << 30 | dy << 31 and the two fractional weights
__device__ __forceinline__ void bilinear_setup(float gx, float gy, int w, int h, uint32_t& off, float& fx, float& fy) {
  const float ix = grid_unnormalize(gx, w);
  const float iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  fx = ix - x0f;
  fy = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const uint32_t dx = x0 + 1 <= w - 1 ? 1u : 0u, dy = y0 + 1 <= h - 1 ? 1u : 0u;
  off = (uint32_t)(y0 * w + x0) | (dx << 30) | (dy << 31);
}

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }

__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// mean over the three pairs of <A,B> / (max(|A|,eps) max(|B|,eps))   (models/matchnerf.py:268-271)
__device__ __forceinline__ float mean_cosine(const float (&q)[9]) {
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    // max(sqrt(x), 1e-8) == sqrt(max(x, 1e-16)): two MUFU.RSQ instead of two IEEE square roots and a division
    acc += q[3 * p] * rsqrtf(fmaxf(q[3 * p + 1], 1e-16f)) * rsqrtf(fmaxf(q[3 * p + 2], 1e-16f));
  }
  return acc * (1.0f / 3.0f);
}

// =====================================================================================================================
// v3: texel-sharing gather.
//
// ncu of v2 (profiles/r01_ncu_summary.md): L1/TEX throughput 88 %, issue 58 %.  In v2 a lane group owns a ray, so every
// one of the 4 rays of a quad pulls its own copy of the 4 x 512 B of a bilinear cell into registers although adjacent
// pixels project into the same cell most of the time (0.25 texel apart at the 1/4 scale, 0.125 at 1/8): the
// L1 -> register path (128 B/clk/SM) carries the same bytes up to four times.  v3 turns the ownership around:
//   * the 32 lanes of a warp own the 32 16-byte slots of a texel (packing v3, pack.cu: lane l holds channels
//     4l..4l+3 of both 128-channel halves), so ONE fully coalesced 512 B load instruction fetches a texel for the warp;
//   * the warp walks the 4 rays of its quad at one depth sample; a bilinear cell (4 texels, 16 registers per lane) is
//     fetched only when the cell changes from one ray to the next (warp-uniform test), then blended per ray with that
//     ray's weights: loads per (view, scale, depth sample) drop from 16 to 4..16 instructions, typically 4..8;
//   * per-sample geometry and tap parameters are computed once with lane = (ray, sample) as in v2, and handed to the
//     gather phase through shared memory (one broadcast 16 B read per use) instead of shuffles; colours are gathered
//     right there by the lane that owns the sample;
//   * the 9 pair products x 4 rays of a lane are reduced with a transposing butterfly (18 + 9 shuffles for the 4-lane
//     fine groups; two more all-reduce steps for the 16-lane coarse groups), after which lane (ray, group) evaluates
//     its own cosine;
//   * conditioning rows are staged in shared memory and written out as 512 B (fp16) / 704 B (fp32) contiguous runs.
constexpr int kWarps3 = 4;          // warps per CTA
constexpr int kStageStride = 28;    // floats per staged row: 22 used, [22, 24) zero, 16-byte aligned rows

__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

__device__ __forceinline__ float dot4(const float2 a0, const float2 a1, const float2 b0, const float2 b1) {
  const float2 s = ffma2(a1, b1, fmul2(a0, b0));
  return s.x + s.y;
}

// One ray sample: this lane's 4 channels of each half of the three views -> its 9 partial products (same order as
// pair_products): pairs (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1)
__device__ __forceinline__ void pair_products4(const uint4& a0, const uint4& a1, const uint4& a2, float (&q)[9]) {
  const float2 v0a0 = __half22float2(h2(a0.x)), v0a1 = __half22float2(h2(a0.y)), v0b0 = __half22float2(h2(a0.z)), v0b1 = __half22float2(h2(a0.w));
  const float2 v1a0 = __half22float2(h2(a1.x)), v1a1 = __half22float2(h2(a1.y)), v1b0 = __half22float2(h2(a1.z)), v1b1 = __half22float2(h2(a1.w));
  const float2 v2a0 = __half22float2(h2(a2.x)), v2a1 = __half22float2(h2(a2.y)), v2b0 = __half22float2(h2(a2.z)), v2b1 = __half22float2(h2(a2.w));
  q[0] = dot4(v0a0, v0a1, v1a0, v1a1); q[1] = dot4(v0a0, v0a1, v0a0, v0a1); q[2] = dot4(v1a0, v1a1, v1a0, v1a1);
  q[3] = dot4(v0b0, v0b1, v2a0, v2a1); q[4] = dot4(v0b0, v0b1, v0b0, v0b1); q[5] = dot4(v2a0, v2a1, v2a0, v2a1);
  q[6] = dot4(v1b0, v1b1, v2b0, v2b1); q[7] = dot4(v1b0, v1b1, v1b0, v1b1); q[8] = dot4(v2b0, v2b1, v2b0, v2b1);
}

__device__ __forceinline__ uint32_t blend1(const __half2 w00, const __half2 w01, const __half2 w10, const __half2 w11,
                                           const uint32_t a, const uint32_t b, const uint32_t c, const uint32_t d) {
  __half2 r = __hmul2(w00, h2(a));
  r = __hfma2(w01, h2(b), r);
  r = __hfma2(w10, h2(c), r);
  r = __hfma2(w11, h2(d), r);
  return *reinterpret_cast<const uint32_t*>(&r);
}

// the 4 texels of a bilinear cell, this lane's 16-byte slot of each.  off = byte offset of texel (y0, x0) inside the
// view's map.  Taps x0+1 / y0+1 are read unconditionally: on the last column / row of a map their weight is exactly 0
// (grid_sample clips the coordinate, so fx = 0 or fy = 0 there) and the address is the next row / the next view / the
// zero tail mnf_pack_features appends (mnf_packed_feature_halves), i.e. finite values in bounds.
__device__ __forceinline__ void load_cell(const char* __restrict__ lane_base, const uint32_t off, const uint32_t rowB, uint4 (&T)[4]) {
  const char* p00 = lane_base + off;
  const char* p10 = p00 + rowB;
  T[0] = __ldg(reinterpret_cast<const uint4*>(p00));
  T[1] = __ldg(reinterpret_cast<const uint4*>(p00 + kFeatCh * 2));
  T[2] = __ldg(reinterpret_cast<const uint4*>(p10));
  T[3] = __ldg(reinterpret_cast<const uint4*>(p10 + kFeatCh * 2));
}

// tap parameters of one (sample, view, scale) for v3: byte offset of texel 00 and the four fp16 weights
__device__ __forceinline__ uint4 make_tap3(float gx, float gy, int w, int h) {
  const float ix = grid_unnormalize(gx, w);
  const float iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = ix - x0f, fy = iy - y0f;
  const __half2 a = __floats2half2_rn((1.f - fx) * (1.f - fy), fx * (1.f - fy));
  const __half2 b = __floats2half2_rn((1.f - fx) * fy, fx * fy);
  return make_uint4((uint32_t)((int)y0f * w + (int)x0f) * (uint32_t)(kFeatCh * 2), *reinterpret_cast<const uint32_t*>(&a),
                    *reinterpret_cast<const uint32_t*>(&b), 0u);
}

// acc + a.lo*b.lo + a.hi*b.hi with fp16 operands and an fp32 accumulator: two FHFMA (the fp16 product is exact in fp32,
// so this equals convert + FFMA bit for bit, without the conversions)
__device__ __forceinline__ float hdot2(const uint32_t a, const uint32_t b, const float acc) {
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
      "mov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %3;\n\t"
      "fma.rn.f32.f16 %0, ah, bh, %0;\n\t}"
      : "=&f"(d) : "r"(a), "r"(b), "f"(acc));
  return d;
}
__device__ __forceinline__ float hdot4(const uint32_t a0, const uint32_t a1, const uint32_t b0, const uint32_t b1) {
  return hdot2(a1, b1, hdot2(a0, b0, 0.f));
}
__device__ __forceinline__ void pair_products4_mixed(const uint4& a0, const uint4& a1, const uint4& a2, float (&q)[9]) {
  q[0] = hdot4(a0.x, a0.y, a1.x, a1.y); q[1] = hdot4(a0.x, a0.y, a0.x, a0.y); q[2] = hdot4(a1.x, a1.y, a1.x, a1.y);
  q[3] = hdot4(a0.z, a0.w, a2.x, a2.y); q[4] = hdot4(a0.z, a0.w, a0.z, a0.w); q[5] = hdot4(a2.x, a2.y, a2.x, a2.y);
  q[6] = hdot4(a1.z, a1.w, a2.z, a2.w); q[7] = hdot4(a1.z, a1.w, a1.z, a1.w); q[8] = hdot4(a2.z, a2.w, a2.z, a2.w);
}

}  // namespace

template <bool kMixed, int kMinBlocks = 4>
__global__ void __launch_bounds__(kWarps3 * 32, kMinBlocks)
gather_cossim_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                     const __half* __restrict__ f0, const int h0, const int w0,
                     const __half* __restrict__ f1, const int h1, const int w1,
                     const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  __shared__ __align__(16) float stage_all[kWarps3][32][kStageStride];   // per warp: 32 samples x 22 values (+ zero pad)
  __shared__ __align__(16) uint4 tp_all[kWarps3][2 * kViews][32];        // per warp: tap parameters [view, scale][sample]
  const uint32_t full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int64_t ray0;                    // index (inside the ray list / range) of the quad's first ray; negative rays are not stored
  int xlimit = kQuad;              // rays of the quad that lie on the image row (tile-list mode)
  if (rays.tile_list) {
    // fix-up pass of the tensor-core gather (gather_tc.cu): 8 blocks x 4 warps = the 32 quads of one listed 16 x 8 pixel tile
    const int slot = blockIdx.x >> 3;
    if (slot >= *rays.tile_count) return;
    const int tile = rays.tile_list[slot];
    const int q = (blockIdx.x & 7) * kWarps3 + wib;
    const int band = rays.band0 + tile / rays.tiles_x, tx = tile - (tile / rays.tiles_x) * rays.tiles_x;
    const int y = band * 8 + (q >> 2), x0 = tx * 16 + (q & 3) * 4;
    if (y >= cams.H || x0 >= cams.W) return;
    ray0 = (int64_t)y * cams.W + x0 - rays.first_ray;
    xlimit = min(kQuad, cams.W - x0);
    if (ray0 + kQuad <= 0) return;
  } else {
    ray0 = ((int64_t)blockIdx.x * kWarps3 + wib) * kQuad;
  }
  if (ray0 >= rays.n_rays) return;
  const int rq = lane >> 3;        // geometry phase: ray of the quad
  const int sj = lane & 7;         // geometry phase: sample inside the chunk
  const int64_t my_ray = max((int64_t)0, min(ray0 + rq, rays.n_rays - 1));   // rays outside the range repeat a valid one (not stored)
  const int64_t pix = rays.points ? 0 : (rays.ray_idx ? rays.ray_idx[my_ray] : rays.first_ray + my_ray);
  float o[3], d[3];
  cast_ray(cams, pix, o, d);
  const int HW = cams.H * cams.W;
  float (*st)[kStageStride] = stage_all[wib];
  uint4 (*tp)[32] = tp_all[wib];
  const char* fb0 = reinterpret_cast<const char*>(f0) + lane * 16;
  const char* fb1 = reinterpret_cast<const char*>(f1) + lane * 16;
  const size_t map0B = (size_t)h0 * w0 * (kFeatCh * 2), map1B = (size_t)h1 * w1 * (kFeatCh * 2);
#pragma unroll
  for (int k = kCond; k < kStageStride; ++k) st[lane][k] = 0.f;
  const bool b0 = lane & 1, b1 = lane & 2;

  for (int s0 = 0; s0 < S; s0 += kChunk) {
    // ------------------------------------------------------------ geometry phase: lane = (ray rq, sample s0 + sj)
    {
      const int s = min(s0 + sj, S - 1);
      float p[3];
      if (rays.points) {                                                         // explicit sample points (query_cond_info)
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __ldg(rays.points + ((size_t)my_ray * S + s) * 3 + i);
      } else {
        const float u = rays.jitter ? rays.jitter[my_ray * S + s] : 0.f;
        const float t = sample_depth(cams, s, S, u);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));   // misc/camera.py:281-286
      }
#pragma unroll
      for (int v = 0; v < kViews; ++v) {
        float uu, vv, zz;
        project_ndc(cams, v, p, uu, vv, zz);
        const float gx = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);                   // matchnerf.py:234
        const float gy = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
        st[lane][19 + v] = (gx > -1.0f && gx < 1.0f && gy > -1.0f && gy < 1.0f) ? 1.f : 0.f;   // :248-250
        tp[v * 2 + 0][lane] = make_tap3(gx, gy, w0, h0);
        tp[v * 2 + 1][lane] = make_tap3(gx, gy, w1, h1);
        // colours (matchnerf.py:245)
        uint32_t coff;
        float fx, fy;
        bilinear_setup(gx, gy, cams.W, cams.H, coff, fx, fy);
        const uint32_t o00 = coff & 0x3fffffffu, dx = (coff >> 30) & 1u, dy = coff >> 31;
        const float4* pr = images + (size_t)v * HW + o00;
        const float4 c00 = __ldg(pr), c01 = __ldg(pr + dx), c10 = __ldg(pr + dy * cams.W), c11 = __ldg(pr + dy * cams.W + dx);
        const float wa0 = (1.f - fx) * (1.f - fy), wb0 = fx * (1.f - fy), wa1 = (1.f - fx) * fy, wb1 = fx * fy;
        st[lane][10 + 3 * v + 0] = (c00.x * wa0 + c01.x * wb0) + (c10.x * wa1 + c11.x * wb1);
        st[lane][10 + 3 * v + 1] = (c00.y * wa0 + c01.y * wb0) + (c10.y * wa1 + c11.y * wb1);
        st[lane][10 + 3 * v + 2] = (c00.z * wa0 + c01.z * wb0) + (c10.z * wa1 + c11.z * wb1);
      }
    }
    __syncwarp();
    // ------------------------------------------------------------ gather phase: the 4 rays of the quad at sample s0 + it
    const int n_it = min(kChunk, S - s0);
    for (int it = 0; it < n_it; ++it) {
#pragma unroll
      for (int sc = 0; sc < 2; ++sc) {
        const char* fb = sc ? fb1 : fb0;
        const uint32_t rowB = (uint32_t)(sc ? w1 : w0) * (uint32_t)(kFeatCh * 2);
        const size_t mapB = sc ? map1B : map0B;
        uint4 T[kViews][4];
        uint32_t cur[kViews];
#pragma unroll
        for (int v = 0; v < kViews; ++v) {        // the first ray's cells of all three views: 12 loads in flight
          cur[v] = tp[v * 2 + sc][it].x;
          load_cell(fb + v * mapB, cur[v], rowB, T[v]);
        }
        uint4 a[kViews][kQuad];
#pragma unroll
        for (int v = 0; v < kViews; ++v) {
#pragma unroll
          for (int i = 0; i < kQuad; ++i) {
            const uint4 t = tp[v * 2 + sc][i * 8 + it];
            if (i > 0 && t.x != cur[v]) {          // warp-uniform: the ray left the cell of its left neighbour
              cur[v] = t.x;
              load_cell(fb + v * mapB, t.x, rowB, T[v]);
            }
            const __half2 w00 = __low2half2(h2(t.y)), w01 = __high2half2(h2(t.y));
            const __half2 w10 = __low2half2(h2(t.z)), w11 = __high2half2(h2(t.z));
            a[v][i].x = blend1(w00, w01, w10, w11, T[v][0].x, T[v][1].x, T[v][2].x, T[v][3].x);
            a[v][i].y = blend1(w00, w01, w10, w11, T[v][0].y, T[v][1].y, T[v][2].y, T[v][3].y);
            a[v][i].z = blend1(w00, w01, w10, w11, T[v][0].z, T[v][1].z, T[v][2].z, T[v][3].z);
            a[v][i].w = blend1(w00, w01, w10, w11, T[v][0].w, T[v][1].w, T[v][2].w, T[v][3].w);
          }
        }
        float q[kQuad][9];
#pragma unroll
        for (int i = 0; i < kQuad; ++i) {
          if (kMixed) pair_products4_mixed(a[0][i], a[1][i], a[2][i], q[i]);
          else pair_products4(a[0][i], a[1][i], a[2][i], q[i]);
        }
        // transposing butterfly over the 4 lanes of a fine group: afterwards lane l holds ray (l & 3)
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float r0 = (b1 ? q[2][k] : q[0][k]) + __shfl_xor_sync(full, b1 ? q[0][k] : q[2][k], 2);
          const float r1 = (b1 ? q[3][k] : q[1][k]) + __shfl_xor_sync(full, b1 ? q[1][k] : q[3][k], 2);
          f[k] = (b0 ? r1 : r0) + __shfl_xor_sync(full, b0 ? r0 : r1, 1);
        }
        const int row = (lane & 3) * 8 + it;
        if (sc == 0) {                              // coarse scale: 64-channel groups = 16 lanes
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            f[k] += __shfl_xor_sync(full, f[k], 4);
            f[k] += __shfl_xor_sync(full, f[k], 8);
          }
          const float sim = mean_cosine(f);
          if ((lane & 12) == 0) st[row][lane >> 4] = sim;
        } else {                                    // fine scale: 16-channel groups = 4 lanes
          st[row][2 + (lane >> 2)] = mean_cosine(f);
        }
      }
    }
    __syncwarp();
    // ------------------------------------------------------------ store: 8 samples of a ray are contiguous in memory
    if (cond_f16) {
#pragma unroll
      for (int i = 0; i < kQuad; ++i) {
        const int sjj = lane >> 2, part = lane & 3;
        if (ray0 + i >= 0 && ray0 + i < rays.n_rays && i < xlimit && sjj < n_it) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (part < 3) {
            const float4 x = *reinterpret_cast<const float4*>(&st[i * 8 + sjj][part * 8]);
            const float4 y = *reinterpret_cast<const float4*>(&st[i * 8 + sjj][part * 8 + 4]);
            const __half2 h0_ = __floats2half2_rn(x.x, x.y), h1_ = __floats2half2_rn(x.z, x.w);
            const __half2 h2_ = __floats2half2_rn(y.x, y.y), h3_ = __floats2half2_rn(y.z, y.w);
            pk.x = *reinterpret_cast<const uint32_t*>(&h0_); pk.y = *reinterpret_cast<const uint32_t*>(&h1_);
            pk.z = *reinterpret_cast<const uint32_t*>(&h2_); pk.w = *reinterpret_cast<const uint32_t*>(&h3_);
          }
          const size_t n = (size_t)(ray0 + i) * S + s0 + sjj;
          *reinterpret_cast<uint4*>(cond_f16 + n * kCondPad + part * 8) = pk;
        }
      }
    }
    if (cond_f32) {
#pragma unroll
      for (int i = 0; i < kQuad; ++i) {
        if (ray0 + i >= 0 && ray0 + i < rays.n_rays && i < xlimit) {
          float* dst = cond_f32 + ((size_t)(ray0 + i) * S + s0) * kCond;      // 22 floats per sample: 8-byte aligned rows
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const int e = 2 * (pass * 32 + lane);
            const int sjj = e / kCond, k = e - sjj * kCond;
            if (sjj < n_it) *reinterpret_cast<float2*>(dst + e) = *reinterpret_cast<const float2*>(&st[i * 8 + sjj][k]);
          }
        }
      }
    }
    __syncwarp();
  }
}

int gather_impl() {
#ifdef MNF_EXPERIMENTS      // -DMNF_EXPERIMENTS builds the mma.sync blend experiment (gather_mma.cu, packing v4); not part of the product library
  static const int impl = [] { const char* e = getenv("MNF_GATHER_IMPL"); return e ? atoi(e) : 3; }();
  return impl == 4 ? 4 : 3;
#else
  return 3;
#endif
}

int launch_gather_mma(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                      const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                      cudaStream_t s);

// gather_tc.cu: 0 = launched (fix = this kernel's tile-list arguments, fix_blocks = its grid), 1 = not applicable, < 0 = error
int launch_gather_tc(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1, int w1,
                     const float* images, float* cond_f32, __half* cond_f16, int* scratch, int scratch_ints, cudaStream_t s,
                     DevRays* fix, int64_t* fix_blocks);

int launch_gather(const DevCams& cams, const DevRays& rays_in, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s, int* scratch, int scratch_ints) {
  if (rays_in.n_rays <= 0) return MNF_OK;
  DevRays rays = rays_in;
  if (cams.local_radius > 0)     // encoder.feature_sample_local_radius > 0: its own (simple) kernel, every kind of ray input
    return launch_gather_local(cams, rays, S, f0, h0, w0, f1, h1, w1, images, cond_f32, cond_f16, s);
#ifdef MNF_EXPERIMENTS
  if (gather_impl() == 4) return launch_gather_mma(cams, rays, S, f0, h0, w0, f1, h1, w1, images, cond_f32, cond_f16, s);
#endif
  if ((int64_t)h0 * w0 >= (1 << 23) || (int64_t)h1 * w1 >= (1 << 23)) { set_error("feature map too large for 32-bit texel offsets"); return MNF_EUNSUPPORTED; }
  const int64_t quads = (rays.n_rays + kQuad - 1) / kQuad;
  int64_t blocks = (quads + kWarps3 - 1) / kWarps3;
  // Contiguous ray ranges (render_by_slices, full images): the tensor-core gather (gather_tc.cu), then THIS kernel in tile-list
  // mode over the tiles whose footprint did not fit its boxes (none in the common case: those blocks exit at once).
  // Explicit ray lists / sample points: this kernel directly.  MNF_GATHER_IMPL=3 forces this kernel for A/B runs.
  static const int tc_on = [] { const char* e = getenv("MNF_GATHER_IMPL"); return e ? atoi(e) != 3 : 1; }();
  if (tc_on && scratch && !rays.ray_idx && !rays.points) {   // every contiguous range, however short: results must not depend on the slicing
    DevRays fix{};
    int64_t fix_blocks = 0;
    const int rc = launch_gather_tc(cams, rays, S, f0, h0, w0, f1, h1, w1, images, cond_f32, cond_f16, scratch, scratch_ints, s, &fix,
                                    &fix_blocks);
    if (rc < 0) return rc;
    if (rc == 0) {
      rays = fix;
      blocks = fix_blocks;
    }
  }
  static const int mixed = [] { const char* e = getenv("MNF_GATHER_MIXED"); return e ? atoi(e) : 1; }();   // A/B knob: FHFMA pair products
  static const int occ = [] { const char* e = getenv("MNF_GATHER_OCC"); return e ? atoi(e) : 5; }();   // CTAs (of 4 warps) per SM: 4 -> 2.44 ms, 5 -> 2.35 ms (96 registers, no spills), 6 -> 2.34 ms per 81,920 rays x 64
  if (mixed && occ == 5)
    gather_cossim_kernel<true, 5><<<(unsigned)blocks, kWarps3 * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                           reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  else if (mixed && occ == 6)
    gather_cossim_kernel<true, 6><<<(unsigned)blocks, kWarps3 * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                           reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  else if (mixed)
    gather_cossim_kernel<true><<<(unsigned)blocks, kWarps3 * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                        reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  else
    gather_cossim_kernel<false><<<(unsigned)blocks, kWarps3 * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                         reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
