// K-gather with encoder.feature_sample_local_radius > 0: every feature sample becomes the MEAN of the bilinear samples at the
// (2r+1)^2 dilated offsets around the projected position.
//
// Replaces sample_features_by_grid's local branch (models/gmflow/utils.py:136-162) inside MatchNeRF.query_cond_info
// (models/matchnerf.py:236-241).  The reference un-normalises the grid with (size-1)/2, adds the integer window offsets times the
// dilation, re-normalises with c' = (size + (2r+1)*dilation - 1)/2 and hands that to F.grid_sample(align_corners=True), which
// un-normalises with (size-1)/2 again -- every sample position is therefore SCALED by (size-1)/(size + (2r+1)*dilation - 1).
// Reproduced as is, operation by operation (same roundings), including that shrink.  Colours and visibility masks use the plain
// grid (matchnerf.py:245-250), exactly as in the radius-0 kernels.
//
// No shipped config sets the option, so this is the simple formulation: one warp per sample, the 32 lanes own the 32 16-byte
// slots of a packed texel (lane l = channels 4l..4l+3 of both 128-channel halves, as in gather.cu / gather_bwd.cu), taps blended
// and averaged in fp32, (dot, |a|^2, |b|^2) reduced over 4 lanes (fine groups) / 16 lanes (coarse groups) with shuffles.
// (2r+1)^2 x 3 views x 2 scales x 4 taps x 512 B of L2 reads per sample.
#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kLocWarps = 8;

struct TapL {
  int off00, dx, dy;            // texel index of tap 00 inside the view's map; +1 steps (0 on the last column / row)
  float w00, w01, w10, w11;
};

__device__ __forceinline__ TapL make_tap_l(float gx, float gy, int w, int h) {
  const float ix = grid_unnormalize(gx, w), iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = ix - x0f, fy = iy - y0f;
  TapL t;
  const int x0 = (int)x0f, y0 = (int)y0f;
  t.off00 = y0 * w + x0;
  t.dx = x0 + 1 <= w - 1 ? 1 : 0;
  t.dy = y0 + 1 <= h - 1 ? w : 0;
  t.w00 = (1.f - fx) * (1.f - fy); t.w01 = fx * (1.f - fy); t.w10 = (1.f - fx) * fy; t.w11 = fx * fy;
  return t;
}

__device__ __forceinline__ void unpack8_l(const uint4 u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}

template <int kLanes>
__device__ __forceinline__ float lane_group_sum(float v) {
#pragma unroll
  for (int off = 1; off < kLanes; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __launch_bounds__(kLocWarps * 32)
gather_cossim_local_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                           const __half* __restrict__ f0, const int h0, const int w0,
                           const __half* __restrict__ f1, const int h1, const int w1,
                           const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  __shared__ float row_all[kLocWarps][kCondPad];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * kLocWarps + wib;
  if (n >= rays.n_rays * (int64_t)S) return;          // warp-uniform
  const int64_t r = n / S;
  const int s = (int)(n - r * S);
  float p[3];
  if (rays.points) {
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __ldg(rays.points + (size_t)n * 3 + i);
  } else {
    const int64_t pix = rays.ray_idx ? rays.ray_idx[r] : rays.first_ray + r;
    float o[3], d[3];
    cast_ray(cams, pix, o, d);
    const float t = sample_depth(cams, s, S, rays.jitter ? rays.jitter[n] : 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));
  }
  float* row = row_all[wib];
  row[lane] = 0.f;
  __syncwarp();
  float gx[kViews], gy[kViews];
  const int HW = cams.H * cams.W;
#pragma unroll
  for (int v = 0; v < kViews; ++v) {
    float uu, vv, zz;
    project_ndc(cams, v, p, uu, vv, zz);
    gx[v] = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);
    gy[v] = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
    if (lane == v) {                                  // colours + mask of view v: plain grid (matchnerf.py:245-250)
      row[19 + v] = (gx[v] > -1.0f && gx[v] < 1.0f && gy[v] > -1.0f && gy[v] < 1.0f) ? 1.f : 0.f;
      const TapL t = make_tap_l(gx[v], gy[v], cams.W, cams.H);
      const float4* pr = images + (size_t)v * HW + t.off00;
      const float4 c00 = __ldg(pr), c01 = __ldg(pr + t.dx), c10 = __ldg(pr + t.dy), c11 = __ldg(pr + t.dy + t.dx);
      row[10 + 3 * v + 0] = (c00.x * t.w00 + c01.x * t.w01) + (c10.x * t.w10 + c11.x * t.w11);
      row[10 + 3 * v + 1] = (c00.y * t.w00 + c01.y * t.w01) + (c10.y * t.w10 + c11.y * t.w11);
      row[10 + 3 * v + 2] = (c00.z * t.w00 + c01.z * t.w01) + (c10.z * t.w10 + c11.z * t.w11);
    }
  }
  const int rad = cams.local_radius, dil = cams.local_dilation;
  const int kwin = 2 * rad + 1;
  const float inv_cnt_den = (float)(kwin * kwin);
#pragma unroll
  for (int sc = 0; sc < 2; ++sc) {
    const __half* fm = sc ? f1 : f0;
    const int h = sc ? h1 : h0, w = sc ? w1 : w0;
    const float cx = __fdiv_rn((float)(w - 1), 2.0f), cy = __fdiv_rn((float)(h - 1), 2.0f);                  // utils.py:141
    const float c2x = __fdiv_rn((float)(w + kwin * dil - 1), 2.0f), c2y = __fdiv_rn((float)(h + kwin * dil - 1), 2.0f);   // :153-154
    float F[kViews][8];
#pragma unroll
    for (int v = 0; v < kViews; ++v) {
#pragma unroll
      for (int i = 0; i < 8; ++i) F[v][i] = 0.f;
      const float ux = __fadd_rn(__fmul_rn(gx[v], cx), cx), uy = __fadd_rn(__fmul_rn(gy[v], cy), cy);          // :142 (not clipped)
      const __half* vbase = fm + (size_t)v * h * w * kFeatCh + lane * 8;
      for (int oy = -rad; oy <= rad; ++oy) {
        const float ny = __fdiv_rn(__fsub_rn(__fadd_rn(uy, (float)(oy * dil)), c2y), c2y);                  // :151, :155
        for (int ox = -rad; ox <= rad; ++ox) {
          const float nx = __fdiv_rn(__fsub_rn(__fadd_rn(ux, (float)(ox * dil)), c2x), c2x);
          const TapL t = make_tap_l(nx, ny, w, h);
          const __half* base = vbase + (size_t)t.off00 * kFeatCh;
          float a[8], b[8], c[8], e[8];
          unpack8_l(__ldg(reinterpret_cast<const uint4*>(base)), a);
          unpack8_l(__ldg(reinterpret_cast<const uint4*>(base + (size_t)t.dx * kFeatCh)), b);
          unpack8_l(__ldg(reinterpret_cast<const uint4*>(base + (size_t)t.dy * kFeatCh)), c);
          unpack8_l(__ldg(reinterpret_cast<const uint4*>(base + (size_t)(t.dy + t.dx) * kFeatCh)), e);
#pragma unroll
          for (int i = 0; i < 8; ++i) F[v][i] += (a[i] * t.w00 + b[i] * t.w01) + (c[i] * t.w10 + e[i] * t.w11);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) F[v][i] = __fdiv_rn(F[v][i], inv_cnt_den);                                // adaptive_avg_pool2d, :160
    }
    // pairs: (v0 half0, v1 half0), (v0 half1, v2 half0), (v1 half1, v2 half1)   (models/matchnerf.py:259-266)
    const float* A[3] = {&F[0][0], &F[0][4], &F[1][4]};
    const float* Bv[3] = {&F[1][0], &F[2][0], &F[2][4]};
    float q[9];
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      float dt = 0.f, aa = 0.f, bb = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { dt += A[pq][i] * Bv[pq][i]; aa += A[pq][i] * A[pq][i]; bb += Bv[pq][i] * Bv[pq][i]; }
      q[3 * pq] = dt; q[3 * pq + 1] = aa; q[3 * pq + 2] = bb;
    }
    float sim = 0.f;
    if (sc == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = lane_group_sum<16>(q[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = lane_group_sum<4>(q[i]);
    }
#pragma unroll
    for (int pq = 0; pq < 3; ++pq)      // <a,b> / (max(|a|, 1e-8) max(|b|, 1e-8)), mean over the pairs (matchnerf.py:268-271)
      sim += q[3 * pq] / (fmaxf(sqrtf(q[3 * pq + 1]), 1e-8f) * fmaxf(sqrtf(q[3 * pq + 2]), 1e-8f));
    sim *= (1.0f / 3.0f);
    if (sc == 0) { if ((lane & 15) == 0) row[lane >> 4] = sim; }
    else         { if ((lane & 3) == 0) row[2 + (lane >> 2)] = sim; }
  }
  __syncwarp();
  if (cond_f32 && lane < kCond) cond_f32[(size_t)n * kCond + lane] = row[lane];
  if (cond_f16 && lane < kCondPad / 2) {
    const __half2 hv = __floats2half2_rn(row[2 * lane], row[2 * lane + 1]);
    reinterpret_cast<__half2*>(cond_f16 + (size_t)n * kCondPad)[lane] = hv;
  }
}

}  // namespace

int launch_gather_local(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1,
                        int w1, const float* images, float* cond_f32, __half* cond_f16, cudaStream_t s) {
  const int64_t n = rays.n_rays * (int64_t)S;
  if (n <= 0) return MNF_OK;
  if (cams.local_radius > 8 || cams.local_dilation < 1) {
    set_error("feature_sample_local_radius = %d / dilation = %d outside [1, 8] / >= 1", cams.local_radius, cams.local_dilation);
    return MNF_EUNSUPPORTED;
  }
  const int64_t grid = (n + kLocWarps - 1) / kLocWarps;
  if (grid > 0x7fffffffLL) { set_error("too many samples for one launch"); return MNF_EINVAL; }
  gather_cossim_local_kernel<<<(unsigned)grid, kLocWarps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                                      reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
