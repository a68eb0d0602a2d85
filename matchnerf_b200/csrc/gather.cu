// K-gather: epipolar projection + bilinear feature gather + grouped cosine similarity.
//
// Replaces MatchNeRF.query_cond_info (models/matchnerf.py:209-293) together with the ray casting /
// depth sampling / projection it depends on (misc/camera.py:255-286, :351-379; matchnerf.py:163-181).
//
// Mapping: one warp owns one ray and walks its S samples.  A lane owns 8 packed channels (16 B) of every
// texel (see pack.cu for the layout), so a 256-channel tap is ONE coalesced 512 B warp load, the three pair
// products are lane-local, and a cosine group is a run of 32/G lanes reduced with xor-shuffles.
// Consecutive rays of an image row sit in consecutive warps of a CTA, so their (nearly identical) epipolar
// footprints are served by L1; the packed DTU maps (39 MB) are L2-resident on B200.
#include "mnf_common.cuh"

namespace mnf {

namespace {

struct Tap {
  int off00, off01, off10, off11;  // element offsets (texel index) of the 4 taps
  float w00, w01, w10, w11;
};

__device__ __forceinline__ Tap make_tap(float gx, float gy, int w, int h) {
  const float ix = grid_unnormalize(gx, w);
  const float iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = ix - x0f, fy = iy - y0f;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
  Tap t;
  t.off00 = y0 * w + x0;
  t.off01 = y0 * w + x1;
  t.off10 = y1 * w + x0;
  t.off11 = y1 * w + x1;
  t.w00 = (1.f - fx) * (1.f - fy);
  t.w01 = fx * (1.f - fy);
  t.w10 = (1.f - fx) * fy;
  t.w11 = fx * fy;
  return t;
}

__device__ __forceinline__ void accum_tap(float acc[8], const uint4 raw, const float wgt) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h2[i]);
    acc[2 * i + 0] = fmaf(wgt, f.x, acc[2 * i + 0]);
    acc[2 * i + 1] = fmaf(wgt, f.y, acc[2 * i + 1]);
  }
}

// Interpolated 8 packed channels of one view at one scale for this lane.
__device__ __forceinline__ void fetch_view(const __half* __restrict__ fmap, const Tap& t, int lane, float acc[8]) {
  const uint4* base = reinterpret_cast<const uint4*>(fmap) + lane;  // 32 uint4 per texel
  const uint4 r00 = __ldg(base + (size_t)t.off00 * 32);
  const uint4 r01 = __ldg(base + (size_t)t.off01 * 32);
  const uint4 r10 = __ldg(base + (size_t)t.off10 * 32);
  const uint4 r11 = __ldg(base + (size_t)t.off11 * 32);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  accum_tap(acc, r00, t.w00);
  accum_tap(acc, r01, t.w01);
  accum_tap(acc, r10, t.w10);
  accum_tap(acc, r11, t.w11);
}

__device__ __forceinline__ float dot4(const float* a, const float* b) {
  return fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])));
}

// Mean over the three view pairs of the grouped cosine similarity; LPG = lanes per group (32 / G).
// Returns the group's value in every lane of that group.  models/matchnerf.py:256-273.
template <int LPG>
__device__ __forceinline__ float pair_cosine(const float a0[8], const float a1[8], const float a2[8]) {
  // halves: [0..3] = half0, [4..7] = half1.  pairs: (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1)
  float q[9];
  q[0] = dot4(a0, a1);          q[1] = dot4(a0, a0);          q[2] = dot4(a1, a1);
  q[3] = dot4(a0 + 4, a2);      q[4] = dot4(a0 + 4, a0 + 4);  q[5] = dot4(a2, a2);
  q[6] = dot4(a1 + 4, a2 + 4);  q[7] = dot4(a1 + 4, a1 + 4);  q[8] = dot4(a2 + 4, a2 + 4);
#pragma unroll
  for (int off = LPG / 2; off >= 1; off >>= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], off);
  }
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const float na = fmaxf(sqrtf(q[3 * p + 1]), 1e-8f);
    const float nb = fmaxf(sqrtf(q[3 * p + 2]), 1e-8f);
    acc += q[3 * p] / (na * nb);
  }
  return acc * (1.0f / 3.0f);
}

}  // namespace

__global__ void __launch_bounds__(256)
gather_cossim_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                     const __half* __restrict__ f0, const int h0, const int w0,
                     const __half* __restrict__ f1, const int h1, const int w1,
                     const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= rays.n_rays) return;
  const int64_t pix = rays.ray_idx ? rays.ray_idx[ray] : rays.first_ray + ray;
  float o[3], d[3];
  cast_ray(cams, pix, o, d);
  const size_t map0 = (size_t)h0 * w0 * kFeatCh, map1 = (size_t)h1 * w1 * kFeatCh;
  const int HW = cams.H * cams.W;

  for (int s = 0; s < S; ++s) {
    const float u = rays.jitter ? rays.jitter[ray * S + s] : 0.f;
    const float t = sample_depth(cams, s, S, u);
    float p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));  // misc/camera.py:281-286

    float gx[kViews], gy[kViews], inside[kViews];
#pragma unroll
    for (int v = 0; v < kViews; ++v) {
      float uu, vv, zz;
      project_ndc(cams, v, p, uu, vv, zz);
      gx[v] = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);  // matchnerf.py:234
      gy[v] = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
      inside[v] = (gx[v] > -1.0f && gx[v] < 1.0f && gy[v] > -1.0f && gy[v] < 1.0f) ? 1.0f : 0.0f;  // :248-250
    }

    float a0[8], a1[8], a2[8];
    // coarse scale, G = 2 -> 16 lanes per group
    fetch_view(f0 + 0 * map0, make_tap(gx[0], gy[0], w0, h0), lane, a0);
    fetch_view(f0 + 1 * map0, make_tap(gx[1], gy[1], w0, h0), lane, a1);
    fetch_view(f0 + 2 * map0, make_tap(gx[2], gy[2], w0, h0), lane, a2);
    const float sim0 = pair_cosine<32 / kG0>(a0, a1, a2);
    // fine scale, G = 8 -> 4 lanes per group
    fetch_view(f1 + 0 * map1, make_tap(gx[0], gy[0], w1, h1), lane, a0);
    fetch_view(f1 + 1 * map1, make_tap(gx[1], gy[1], w1, h1), lane, a1);
    fetch_view(f1 + 2 * map1, make_tap(gx[2], gy[2], w1, h1), lane, a2);
    const float sim1 = pair_cosine<32 / kG1>(a0, a1, a2);

    // colours: lanes 0..11 = (view, tap); matchnerf.py:245
    float3 col = make_float3(0.f, 0.f, 0.f);
    {
      const int v = min(lane >> 2, kViews - 1), tap = lane & 3;
      const float gxv = v == 0 ? gx[0] : (v == 1 ? gx[1] : gx[2]);
      const float gyv = v == 0 ? gy[0] : (v == 1 ? gy[1] : gy[2]);
      const Tap tp = make_tap(gxv, gyv, cams.W, cams.H);
      const int off = tap == 0 ? tp.off00 : (tap == 1 ? tp.off01 : (tap == 2 ? tp.off10 : tp.off11));
      const float wt = tap == 0 ? tp.w00 : (tap == 1 ? tp.w01 : (tap == 2 ? tp.w10 : tp.w11));
      if (lane < 4 * kViews) {
        const float4 c = __ldg(images + (size_t)v * HW + off);
        col = make_float3(c.x * wt, c.y * wt, c.z * wt);
      }
#pragma unroll
      for (int off2 = 1; off2 <= 2; off2 <<= 1) {
        col.x += __shfl_xor_sync(0xffffffffu, col.x, off2);
        col.y += __shfl_xor_sync(0xffffffffu, col.y, off2);
        col.z += __shfl_xor_sync(0xffffffffu, col.z, off2);
      }
    }

    // route value k to lane k: [0,2) coarse sims, [2,10) fine sims, [10,19) colours, [19,22) masks
    const int k = lane;
    const float s0 = __shfl_sync(0xffffffffu, sim0, (k & 1) * 16);
    const float s1 = __shfl_sync(0xffffffffu, sim1, ((k - 2) & 7) * 4);
    const int cv = (k >= 10 && k < 19) ? (k - 10) / 3 : 0;
    const float cr = __shfl_sync(0xffffffffu, col.x, cv * 4);
    const float cg = __shfl_sync(0xffffffffu, col.y, cv * 4);
    const float cb = __shfl_sync(0xffffffffu, col.z, cv * 4);
    float val = 0.f;
    if (k < 2) val = s0;
    else if (k < 10) val = s1;
    else if (k < 19) { const int c = (k - 10) % 3; val = c == 0 ? cr : (c == 1 ? cg : cb); }
    else if (k < 22) val = k == 19 ? inside[0] : (k == 20 ? inside[1] : inside[2]);
    const size_t n = (size_t)ray * S + s;
    if (cond_f32 && k < kCond) cond_f32[n * kCond + k] = val;
    if (cond_f16) cond_f16[n * kCondPad + k] = __float2half_rn(val);
  }
}

int launch_gather(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  const int warps = 8;
  const int64_t blocks = (rays.n_rays + warps - 1) / warps;
  gather_cossim_kernel<<<(unsigned)blocks, warps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                              reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
