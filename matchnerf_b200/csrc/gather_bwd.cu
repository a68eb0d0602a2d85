// K-gather backward: d(loss)/d(cosine similarities) -> d(loss)/d(feature maps), the training-step half of K-gather.
//
// Reverses MatchNeRF.query_cond_info (models/matchnerf.py:209-293) for the quantities that carry gradient to parameters: the ten
// grouped cosine similarities of a sample (feat_info) as functions of the bilinearly sampled encoder features
// (F.grid_sample backward = a scatter-add of the blended-feature gradient with the four tap weights; nn.CosineSimilarity
// backward with its eps clamp).  Colours and visibility masks depend on the input images and the cameras only (no parameters),
// and the sample positions do not depend on the feature maps, so nothing else flows (the reference's autograd graph has the
// same leaves).
//
// One warp per sample.  The 32 lanes own the 32 16-byte slots of a packed texel exactly as in the forward kernel (gather.cu, packing
// v3: lane l holds channels 4l..4l+3 of BOTH 128-channel halves), so the three pair products are lane-local, a fine cosine
// group (16 channels) is 4 lanes and a coarse one (64 channels) 16 lanes.  The warp re-derives the sample's geometry with the
// forward kernel's arithmetic (same taps), blends the three views in fp32, reduces (dot, |a|^2, |b|^2) per pair and group with
// shuffles, forms dF and scatters it with `red.global.add.v4.f32` into channels-last fp32 gradient maps in the PACKED channel
// order (the host un-permutes them once per step).  HBM/L2-atomic bound: 3 views x 2 scales x 4 taps x 1 KB per sample.
#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kBwdWarps = 8;

struct TapB {
  int x0, y0, dx, dy;
  float w00, w01, w10, w11;
};

__device__ __forceinline__ TapB make_tap_b(float gx, float gy, int w, int h) {
  const float ix = grid_unnormalize(gx, w), iy = grid_unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float fx = ix - x0f, fy = iy - y0f;
  TapB t;
  t.x0 = (int)x0f; t.y0 = (int)y0f;
  t.dx = t.x0 + 1 <= w - 1 ? 1 : 0; t.dy = t.y0 + 1 <= h - 1 ? 1 : 0;      // (on the last column / row fx / fy is exactly 0)
  t.w00 = (1.f - fx) * (1.f - fy); t.w01 = fx * (1.f - fy); t.w10 = (1.f - fx) * fy; t.w11 = fx * fy;
  return t;
}

__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int kLanes>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = 1; off < kLanes; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// d cos(a, b) with the eps clamp of nn.CosineSimilarity (cos = <a,b> / (max(|a|, eps) max(|b|, eps)), eps = 1e-8):
// returns the coefficients (ca_b, ca_a, cb_a, cb_b) such that da = ca_b * b - ca_a * a and db = cb_a * a - cb_b * b
__device__ __forceinline__ void cos_grad_coeffs(float dot, float aa, float bb, float dcos, float& ca_b, float& ca_a, float& cb_a, float& cb_b) {
  const float ia = rsqrtf(fmaxf(aa, 1e-16f)), ib = rsqrtf(fmaxf(bb, 1e-16f));
  const float iab = ia * ib * dcos;
  ca_b = iab;
  cb_a = iab;
  ca_a = aa > 1e-16f ? dot * iab * ia * ia : 0.f;
  cb_b = bb > 1e-16f ? dot * iab * ib * ib : 0.f;
}

__global__ void __launch_bounds__(kBwdWarps * 32)
gather_cossim_bwd_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                         const __half* __restrict__ f0, const int h0, const int w0, const __half* __restrict__ f1, const int h1,
                         const int w1, const float* __restrict__ dcond, float* __restrict__ g0, float* __restrict__ g1) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * kBwdWarps + (threadIdx.x >> 5);
  if (n >= rays.n_rays * (int64_t)S) return;
  const int64_t r = n / S;
  const int s = (int)(n - r * S);
  // the ten cosine gradients of this sample (lanes 0..9), zero rows are skipped
  const float dc_mine = lane < 10 ? dcond[n * kCond + lane] : 0.f;
  if (__ballot_sync(0xffffffffu, dc_mine != 0.f) == 0u) return;
  const int64_t pix = rays.ray_idx ? rays.ray_idx[r] : rays.first_ray + r;
  float o[3], d[3];
  cast_ray(cams, pix, o, d);
  const float u = rays.jitter ? rays.jitter[n] : 0.f;
  const float t = sample_depth(cams, s, S, u);
  float p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));
  float gx[kViews], gy[kViews];
#pragma unroll
  for (int v = 0; v < kViews; ++v) {
    float uu, vv, zz;
    project_ndc(cams, v, p, uu, vv, zz);
    gx[v] = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);
    gy[v] = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
  }
#pragma unroll
  for (int sc = 0; sc < 2; ++sc) {
    const __half* fm = sc ? f1 : f0;
    float* gm = sc ? g1 : g0;
    const int h = sc ? h1 : h0, w = sc ? w1 : w0;
    TapB tap[kViews];
    float F[kViews][8];                       // blended features: [view][half 0: 4 channels | half 1: 4 channels]
#pragma unroll
    for (int v = 0; v < kViews; ++v) {
      tap[v] = make_tap_b(gx[v], gy[v], w, h);
      const __half* base = fm + ((size_t)v * h * w + (size_t)tap[v].y0 * w + tap[v].x0) * kFeatCh + lane * 8;
      float a[8], b[8], c[8], e[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(base)), a);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)tap[v].dx * kFeatCh)), b);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (size_t)tap[v].dy * w * kFeatCh)), c);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + ((size_t)tap[v].dy * w + tap[v].dx) * kFeatCh)), e);
#pragma unroll
      for (int i = 0; i < 8; ++i) F[v][i] = (a[i] * tap[v].w00 + b[i] * tap[v].w01) + (c[i] * tap[v].w10 + e[i] * tap[v].w11);
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
    float dcos;
    if (sc == 0) {                            // coarse: 2 groups of 64 channels = 16 lanes each; cond[0..1]
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = group_sum<16>(q[i]);
      dcos = __shfl_sync(0xffffffffu, dc_mine, lane >> 4);
    } else {                                  // fine: 8 groups of 16 channels = 4 lanes each; cond[2..9]
#pragma unroll
      for (int i = 0; i < 9; ++i) q[i] = group_sum<4>(q[i]);
      dcos = __shfl_sync(0xffffffffu, dc_mine, 2 + (lane >> 2));
    }
    dcos *= (1.0f / 3.0f);                    // mean over the three pairs
    float dF[kViews][8];
#pragma unroll
    for (int v = 0; v < kViews; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) dF[v][i] = 0.f;
    float* dA[3] = {&dF[0][0], &dF[0][4], &dF[1][4]};
    float* dB[3] = {&dF[1][0], &dF[2][0], &dF[2][4]};
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      float ca_b, ca_a, cb_a, cb_b;
      cos_grad_coeffs(q[3 * pq], q[3 * pq + 1], q[3 * pq + 2], dcos, ca_b, ca_a, cb_a, cb_b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dA[pq][i] += ca_b * Bv[pq][i] - ca_a * A[pq][i];
        dB[pq][i] += cb_a * A[pq][i] - cb_b * Bv[pq][i];
      }
    }
    if (dcos != 0.f) {
#pragma unroll
      for (int v = 0; v < kViews; ++v) {
        float* base = gm + ((size_t)v * h * w + (size_t)tap[v].y0 * w + tap[v].x0) * kFeatCh + lane * 8;
        const float ws[4] = {tap[v].w00, tap[v].w01, tap[v].w10, tap[v].w11};
        const size_t offs[4] = {0, (size_t)tap[v].dx * kFeatCh, (size_t)tap[v].dy * w * kFeatCh, ((size_t)tap[v].dy * w + tap[v].dx) * kFeatCh};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (ws[k] != 0.f) {
            float* pk = base + offs[k];
            red_add_v4(pk, ws[k] * dF[v][0], ws[k] * dF[v][1], ws[k] * dF[v][2], ws[k] * dF[v][3]);
            red_add_v4(pk + 4, ws[k] * dF[v][4], ws[k] * dF[v][5], ws[k] * dF[v][6], ws[k] * dF[v][7]);
          }
        }
      }
    }
  }
}

}  // namespace

int launch_gather_bwd(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1,
                      int w1, const float* dcond, float* g0, float* g1, cudaStream_t s) {
  const int64_t n = rays.n_rays * (int64_t)S;
  if (n <= 0) return MNF_OK;
  const unsigned grid = (unsigned)((n + kBwdWarps - 1) / kBwdWarps);
  gather_cossim_bwd_kernel<<<grid, kBwdWarps * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1, dcond, g0, g1);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
