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
gather_cossim_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
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

int launch_gather(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  const int warps = 8;
  const int64_t quads = (rays.n_rays + kQuad - 1) / kQuad;
  const int64_t blocks = (quads + warps - 1) / warps;
  static const int occ = [] { const char* e = getenv("MNF_GATHER_OCC"); return e ? atoi(e) : 2; }();   // A/B knob: CTAs per SM
  if (occ == 3)
    gather_cossim_kernel<3><<<(unsigned)blocks, warps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                   reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  else
    gather_cossim_kernel<2><<<(unsigned)blocks, warps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                   reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
