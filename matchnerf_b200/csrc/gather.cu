// K-gather: epipolar projection + bilinear feature gather + grouped cosine similarity.
//
// Replaces MatchNeRF.query_cond_info (models/matchnerf.py:209-293) together with the ray casting /
// depth sampling / projection it depends on (misc/camera.py:255-286, :351-379; matchnerf.py:163-181).
//
// Work decomposition (v2; ncu of v1 showed the kernel issue-bound at 1442 warp-instructions per sample, most of
// them per-sample scalar geometry replicated in all 32 lanes):
//   * a warp owns a QUAD of 4 consecutive rays and walks their samples in lock step;
//   * geometry phase, lane = (ray of the quad, sample of an 8-sample chunk): projection into the 3 source views and
//     the bilinear tap set-up are computed ONCE per sample, 32 samples at a time, and kept in registers;
//   * gather phase, 8 iterations per chunk: the 8 lanes of group g handle the sample of ray g; lane l owns 32 packed
//     channels (64 B) of every texel, receives its sample's tap parameters by shuffle, blends the 4 taps with packed
//     half2 FMAs, and accumulates the 9 pair products (3 dots + 6 squared norms) with packed fp32 FMAs (fma.f32x2).
// The packed feature layout (pack.cu) makes the three pair products lane-local, a fine-scale cosine group (16 ch)
// lane-local and a coarse group (64 ch) a run of 4 lanes (2 xor-shuffles).  The 4 rays of a quad are adjacent pixels:
// their taps mostly coincide, so the 4 groups of one load instruction hit the same L1 lines.
#include <cstdlib>

#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kQuad = 4;        // rays per warp
constexpr int kChunk = 8;       // samples per geometry phase (kQuad * kChunk == 32 lanes)

struct TapParam {               // one (view, scale): 3 registers
  uint32_t off;                 // texel index of tap 00 | dx << 30 | dy << 31
  uint32_t w0;                  // half2 (w00, w01)
  uint32_t w1;                  // half2 (w10, w11)
};

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

__device__ __forceinline__ TapParam make_tap(float gx, float gy, int w, int h) {
  TapParam t;
  float fx, fy;
  bilinear_setup(gx, gy, w, h, t.off, fx, fy);
  const __half2 a = __floats2half2_rn((1.f - fx) * (1.f - fy), fx * (1.f - fy));
  const __half2 b = __floats2half2_rn((1.f - fx) * fy, fx * fy);
  t.w0 = *reinterpret_cast<const uint32_t*>(&a);
  t.w1 = *reinterpret_cast<const uint32_t*>(&b);
  return t;
}

__device__ __forceinline__ TapParam shfl_tap(const TapParam& t, int src) {
  TapParam r;
  r.off = __shfl_sync(0xffffffffu, t.off, src);
  r.w0 = __shfl_sync(0xffffffffu, t.w0, src);
  r.w1 = __shfl_sync(0xffffffffu, t.w1, src);
  return r;
}

__device__ __forceinline__ __half2 h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }

// blend 4 taps of this lane's 32 packed channels: acc[16] half2
__device__ __forceinline__ void fetch_blend(const __half* __restrict__ fmap, int w, const TapParam& t, int li, __half2 (&acc)[16]) {
  const uint32_t o00 = t.off & 0x3fffffffu;
  const uint32_t dx = (t.off >> 30) & 1u, dy = t.off >> 31;
  // 32 uint4 per texel; load j of lane li is uint4 number 8*j + li, so the 8 lanes of a group read 128 contiguous bytes
  const uint4* p00 = reinterpret_cast<const uint4*>(fmap) + (size_t)o00 * 32 + li;
  const uint4* p01 = p00 + dx * 32;
  const uint4* p10 = p00 + (size_t)dy * w * 32;
  const uint4* p11 = p10 + dx * 32;
  const __half2 w00 = __low2half2(h2(t.w0)), w01 = __high2half2(h2(t.w0));
  const __half2 w10 = __low2half2(h2(t.w1)), w11 = __high2half2(h2(t.w1));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 a = __ldg(p00 + 8 * j), b = __ldg(p01 + 8 * j), c = __ldg(p10 + 8 * j), d = __ldg(p11 + 8 * j);
    const uint32_t* av = &a.x; const uint32_t* bv = &b.x; const uint32_t* cv = &c.x; const uint32_t* dv = &d.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 r = __hmul2(w00, h2(av[i]));
      r = __hfma2(w01, h2(bv[i]), r);
      r = __hfma2(w10, h2(cv[i]), r);
      r = __hfma2(w11, h2(dv[i]), r);
      acc[j * 4 + i] = r;
    }
  }
}

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

// 16-channel dot products of this lane for the three view pairs: q[3p] = <A,B>, q[3p+1] = <A,A>, q[3p+2] = <B,B>
// halves: acc[0..7] = half0 channels 16l..16l+15, acc[8..15] = half1 channels 16l..16l+15 (loads j=0,1 / j=2,3).  pairs (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1)
__device__ __forceinline__ void pair_products(const __half2 (&a0)[16], const __half2 (&a1)[16], const __half2 (&a2)[16], float (&q)[9]) {
  float2 s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 v0a = __half22float2(a0[i]), v0b = __half22float2(a0[8 + i]);
    const float2 v1a = __half22float2(a1[i]), v1b = __half22float2(a1[8 + i]);
    const float2 v2a = __half22float2(a2[i]), v2b = __half22float2(a2[8 + i]);
    s[0] = ffma2(v0a, v1a, s[0]); s[1] = ffma2(v0a, v0a, s[1]); s[2] = ffma2(v1a, v1a, s[2]);
    s[3] = ffma2(v0b, v2a, s[3]); s[4] = ffma2(v0b, v0b, s[4]); s[5] = ffma2(v2a, v2a, s[5]);
    s[6] = ffma2(v1b, v2b, s[6]); s[7] = ffma2(v1b, v1b, s[7]); s[8] = ffma2(v2b, v2b, s[8]);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) q[i] = s[i].x + s[i].y;
}

// mean over the three pairs of <A,B> / (max(|A|,eps) max(|B|,eps))   (models/matchnerf.py:268-271)
__device__ __forceinline__ float mean_cosine(const float (&q)[9]) {
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const float na = fmaxf(sqrtf(q[3 * p + 1]), 1e-8f);
    const float nb = fmaxf(sqrtf(q[3 * p + 2]), 1e-8f);
    acc += __fdividef(q[3 * p], na * nb);
  }
  return acc * (1.0f / 3.0f);
}

}  // namespace

template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
gather_cossim_v2_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                     const __half* __restrict__ f0, const int h0, const int w0,
                     const __half* __restrict__ f1, const int h1, const int w1,
                     const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  __shared__ __align__(16) float stage[8][kQuad][kCondPad];   // per warp: 4 samples x 32 values
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t quad = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t ray0 = quad * kQuad;
  if (ray0 >= rays.n_rays) return;
  const int rq = lane >> 3;        // geometry phase: ray of the quad;   gather phase: group id == ray of the quad
  const int sj = lane & 7;         // geometry phase: sample inside the chunk; gather phase: lane inside the group
  const int li = sj;
  const int64_t my_ray = min(ray0 + rq, rays.n_rays - 1);       // rays past the end repeat the last one (not stored)
  const bool ray_ok = ray0 + rq < rays.n_rays;
  const int64_t pix = rays.ray_idx ? rays.ray_idx[my_ray] : rays.first_ray + my_ray;
  float o[3], d[3];
  cast_ray(cams, pix, o, d);
  const size_t map0 = (size_t)h0 * w0 * kFeatCh, map1 = (size_t)h1 * w1 * kFeatCh;
  const int HW = cams.H * cams.W;
  float (*st)[kCondPad] = stage[wib];

  for (int s0 = 0; s0 < S; s0 += kChunk) {
    // ------------------------------------------------------------ geometry phase: lane = (ray rq, sample s0 + sj)
    TapParam tc0[kViews], tc1[kViews];
    uint32_t coff[kViews];
    float cfx[kViews], cfy[kViews];
    uint32_t inside_bits = 0;
    {
      const int s = min(s0 + sj, S - 1);
      const float u = rays.jitter ? rays.jitter[my_ray * S + s] : 0.f;
      const float t = sample_depth(cams, s, S, u);
      float p[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));   // misc/camera.py:281-286
#pragma unroll
      for (int v = 0; v < kViews; ++v) {
        float uu, vv, zz;
        project_ndc(cams, v, p, uu, vv, zz);
        const float gx = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);                   // matchnerf.py:234
        const float gy = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
        if (gx > -1.0f && gx < 1.0f && gy > -1.0f && gy < 1.0f) inside_bits |= 1u << v;   // :248-250
        tc0[v] = make_tap(gx, gy, w0, h0);
        tc1[v] = make_tap(gx, gy, w1, h1);
        bilinear_setup(gx, gy, cams.W, cams.H, coff[v], cfx[v], cfy[v]);
      }
    }
    // ------------------------------------------------------------ gather phase: group rq handles sample (ray rq, s0 + it)
    const int n_it = min(kChunk, S - s0);
    for (int it = 0; it < n_it; ++it) {
      const int src = (lane & 24) | it;      // lane of the geometry phase that holds this group's sample
      __half2 a0[16], a1[16], a2[16];
      float q[9];
      // coarse scale: 64-channel groups = 4 lanes
      fetch_blend(f0 + 0 * map0, w0, shfl_tap(tc0[0], src), li, a0);
      fetch_blend(f0 + 1 * map0, w0, shfl_tap(tc0[1], src), li, a1);
      fetch_blend(f0 + 2 * map0, w0, shfl_tap(tc0[2], src), li, a2);
      pair_products(a0, a1, a2, q);
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        q[i] += __shfl_xor_sync(0xffffffffu, q[i], 1);
        q[i] += __shfl_xor_sync(0xffffffffu, q[i], 2);
      }
      const float sim0 = mean_cosine(q);      // lanes 0-3: group 0, lanes 4-7: group 1
      // fine scale: 16-channel groups are lane-local
      fetch_blend(f1 + 0 * map1, w1, shfl_tap(tc1[0], src), li, a0);
      fetch_blend(f1 + 1 * map1, w1, shfl_tap(tc1[1], src), li, a1);
      fetch_blend(f1 + 2 * map1, w1, shfl_tap(tc1[2], src), li, a2);
      pair_products(a0, a1, a2, q);
      const float sim1 = mean_cosine(q);      // lane li: group li

      // colours (matchnerf.py:245): lanes 0..5 of the group = (view, tap row)
      float3 col = make_float3(0.f, 0.f, 0.f);
      {
        const int v = min(li >> 1, kViews - 1), rowsel = li & 1;
        // (the view is chosen by the RECEIVING lane, so all three are shuffled and selected afterwards)
        const uint32_t o0 = __shfl_sync(0xffffffffu, coff[0], src), o1 = __shfl_sync(0xffffffffu, coff[1], src), o2 = __shfl_sync(0xffffffffu, coff[2], src);
        const float x0 = __shfl_sync(0xffffffffu, cfx[0], src), x1 = __shfl_sync(0xffffffffu, cfx[1], src), x2 = __shfl_sync(0xffffffffu, cfx[2], src);
        const float y0 = __shfl_sync(0xffffffffu, cfy[0], src), y1 = __shfl_sync(0xffffffffu, cfy[1], src), y2 = __shfl_sync(0xffffffffu, cfy[2], src);
        const uint32_t offv = v == 0 ? o0 : (v == 1 ? o1 : o2);
        const float fx = v == 0 ? x0 : (v == 1 ? x1 : x2);
        const float fy = v == 0 ? y0 : (v == 1 ? y1 : y2);
        if (li < 2 * kViews) {
          const uint32_t o00 = offv & 0x3fffffffu, dx = (offv >> 30) & 1u, dy = offv >> 31;
          const float4* pr = images + (size_t)v * HW + o00 + (rowsel ? dy * cams.W : 0u);
          const float4 c0 = __ldg(pr), c1 = __ldg(pr + dx);
          const float wy = rowsel ? fy : 1.f - fy;
          const float wa = (1.f - fx) * wy, wb = fx * wy;
          col = make_float3(c0.x * wa + c1.x * wb, c0.y * wa + c1.y * wb, c0.z * wa + c1.z * wb);
        }
        col.x += __shfl_xor_sync(0xffffffffu, col.x, 1);
        col.y += __shfl_xor_sync(0xffffffffu, col.y, 1);
        col.z += __shfl_xor_sync(0xffffffffu, col.z, 1);
      }
      const uint32_t inb = __shfl_sync(0xffffffffu, inside_bits, src);

      // stage the 32 values of this group's sample, then store 4 samples with coalesced writes
      float* row = st[rq];
      if (li == 0) row[0] = sim0;
      if (li == 4) row[1] = sim0;
      row[2 + li] = sim1;
      if (li < 6 && !(li & 1)) { row[10 + 3 * (li >> 1)] = col.x; row[11 + 3 * (li >> 1)] = col.y; row[12 + 3 * (li >> 1)] = col.z; }
      if (li == 7) {
        row[19] = (inb & 1u) ? 1.f : 0.f; row[20] = (inb & 2u) ? 1.f : 0.f; row[21] = (inb & 4u) ? 1.f : 0.f;
#pragma unroll
        for (int k = 22; k < 32; ++k) row[k] = 0.f;
      }
      __syncwarp();
      {
        const int64_t r_out = ray0 + rq;
        const size_t n = (size_t)my_ray * S + (s0 + it);
        const float4 v4 = *reinterpret_cast<const float4*>(&st[rq][li * 4]);
        if (r_out < rays.n_rays) {
          if (cond_f16) {
            const __half2 lo = __floats2half2_rn(v4.x, v4.y), hi = __floats2half2_rn(v4.z, v4.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(cond_f16 + n * kCondPad + li * 4) = pk;
          }
          if (cond_f32) {
            float* dst = cond_f32 + n * kCond + li * 4;       // 22 floats per sample: 88 B rows are only 8 B aligned
            if (li < 5) { *reinterpret_cast<float2*>(dst) = make_float2(v4.x, v4.y); *reinterpret_cast<float2*>(dst + 2) = make_float2(v4.z, v4.w); }
            else if (li == 5) { *reinterpret_cast<float2*>(dst) = make_float2(v4.x, v4.y); }
          }
        }
      }
      __syncwarp();
    }
  }
  (void)ray_ok;
}

// =====================================================================================================================
// v3: texel-sharing gather (default).
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
namespace {

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
  const int64_t quad = (int64_t)blockIdx.x * kWarps3 + wib;
  const int64_t ray0 = quad * kQuad;
  if (ray0 >= rays.n_rays) return;
  const int rq = lane >> 3;        // geometry phase: ray of the quad
  const int sj = lane & 7;         // geometry phase: sample inside the chunk
  const int64_t my_ray = min(ray0 + rq, rays.n_rays - 1);       // rays past the end repeat the last one (not stored)
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
        if (ray0 + i < rays.n_rays && sjj < n_it) {
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
        if (ray0 + i < rays.n_rays) {
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
  static const int impl = [] { const char* e = getenv("MNF_GATHER_IMPL"); return e ? atoi(e) : 3; }();   // A/B knob: 2 = v2, 3 = v3, 4 = v4 (tensor-core blend)
  return (impl == 2 || impl == 4) ? impl : 3;
}

int launch_gather_mma(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                      const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                      cudaStream_t s);

static int launch_gather_v2(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  const int warps = 8;
  const int64_t quads = (rays.n_rays + kQuad - 1) / kQuad;
  const int64_t blocks = (quads + warps - 1) / warps;
  static const int occ = [] { const char* e = getenv("MNF_GATHER_OCC"); return e ? atoi(e) : 2; }();   // A/B knob: CTAs per SM
  if (occ == 3)
    gather_cossim_v2_kernel<3><<<(unsigned)blocks, warps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                   reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  else
    gather_cossim_v2_kernel<2><<<(unsigned)blocks, warps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                   reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

int launch_gather(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  if (gather_impl() == 4) return launch_gather_mma(cams, rays, S, f0, h0, w0, f1, h1, w1, images, cond_f32, cond_f16, s);
  if (gather_impl() == 2 && rays.points) { set_error("explicit sample points need the v3 gather kernel (unset MNF_GATHER_IMPL)"); return MNF_EUNSUPPORTED; }
  if (gather_impl() == 2) return launch_gather_v2(cams, rays, S, f0, h0, w0, f1, h1, w1, images, cond_f32, cond_f16, s);
  const int64_t quads = (rays.n_rays + kQuad - 1) / kQuad;
  const int64_t blocks = (quads + kWarps3 - 1) / kWarps3;
  if ((int64_t)h0 * w0 >= (1 << 23) || (int64_t)h1 * w1 >= (1 << 23)) { set_error("feature map too large for 32-bit texel offsets"); return MNF_EUNSUPPORTED; }
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
