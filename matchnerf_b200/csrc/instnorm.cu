// Fused InstanceNorm2d (no affine, eps 1e-5, biased variance) + ReLU + residual add for the GMFlow CNN backbone
// (models/gmflow/backbone.py:6-36, :101-122): the reference runs F.instance_norm (two batch-norm kernels), a ReLU, an add
// and another ReLU as separate passes over [N, C, H, W] fp32 activations (ncu launch list: 2.65 of the backbone's ~5 ms
// under the profiler).  One kernel does
//     y = IN(x)                       mode 0   (downsample branch: conv1x1 -> norm)
//     y = relu(IN(x))                 mode 1
//     y = relu(res + relu(IN(x)))     mode 2   (block output: relu(skip + y))
// One CTA per (n, c) plane, contiguous NCHW.  Pass 1 reduces sum / sum of squares (fp32 per thread, double across the
// block), pass 2 re-reads the plane (L2-resident: the largest plane is 320 KB), normalises and writes.  HBM/L2-bound:
// 8 B read + 4 B written per element (+ 4 B for the residual) instead of ~28 B over five kernels.
#include "mnf_common.cuh"

namespace mnf {

namespace {
constexpr int kInThreads = 512;
}

__global__ void __launch_bounds__(kInThreads, 2)
instance_norm_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y, const int hw,
                     const int mode, const float eps) {
  __shared__ double red[2][kInThreads / 32];
  __shared__ float stats[2];
  const size_t plane = (size_t)blockIdx.x * hw;
  const float* xp = x + plane;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n4 = (hw & 3) == 0 && ((reinterpret_cast<uintptr_t>(xp) & 15) == 0) ? hw >> 2 : 0;   // vector part
  float s = 0.f, ss = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(xp);
  for (int i = tid; i < n4; i += kInThreads) {
    const float4 v = __ldg(x4 + i);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  for (int i = n4 * 4 + tid; i < hw; i += kInThreads) {
    const float v = __ldg(xp + i);
    s += v;
    ss += v * v;
  }
  double ds = s, dss = ss;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    ds += __shfl_xor_sync(0xffffffffu, ds, off);
    dss += __shfl_xor_sync(0xffffffffu, dss, off);
  }
  if (lane == 0) { red[0][wid] = ds; red[1][wid] = dss; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < kInThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    const double mean = a / hw;
    const double var = fmax(b / hw - mean * mean, 0.0);
    stats[0] = (float)mean;
    stats[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = stats[0], rstd = stats[1];
  float* yp = y + plane;
  const float* rp = res ? res + plane : nullptr;
  auto f = [&](float v, float r) {
    float o = (v - mean) * rstd;
    if (mode >= 1) o = fmaxf(o, 0.f);
    if (mode == 2) o = fmaxf(o + r, 0.f);
    return o;
  };
  const bool vec_ok = n4 > 0 && ((reinterpret_cast<uintptr_t>(yp) & 15) == 0) && (!rp || (reinterpret_cast<uintptr_t>(rp) & 15) == 0);
  const int nv = vec_ok ? n4 : 0;
  for (int i = tid; i < nv; i += kInThreads) {
    const float4 v = __ldg(x4 + i);
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rp) r = __ldg(reinterpret_cast<const float4*>(rp) + i);
    reinterpret_cast<float4*>(yp)[i] = make_float4(f(v.x, r.x), f(v.y, r.y), f(v.z, r.z), f(v.w, r.w));
  }
  for (int i = nv * 4 + tid; i < hw; i += kInThreads) yp[i] = f(__ldg(xp + i), rp ? __ldg(rp + i) : 0.f);
}

int launch_instance_norm(const float* x, const float* res, float* y, int64_t planes, int hw, int mode, float eps, cudaStream_t s) {
  if (planes <= 0 || hw <= 0) return MNF_OK;
  instance_norm_kernel<<<(unsigned)planes, kInThreads, 0, s>>>(x, res, y, hw, mode, eps);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf

// ---------------------------------------------------------------------------------------------------------------------
// Channels-last (NHWC) variant, fp32 or fp16: cuDNN's NHWC convolutions need no layout conversion kernels around them (backbone
// convolutions at DTU size: 1.24 ms fp32 NCHW -> 0.74 ms fp32 NHWC (TF32) -> 0.46 ms fp16 NHWC).
// x, res, y: [N][HW][C], C a multiple of 8.  Two launches: per-(n, c) sum / sum of squares (fp32 per thread over <= a few
// hundred pixels, fixed-order reduction inside the CTA, per-CTA partials reduced by the image's last CTA), then normalise +
// ReLU / residual with 16-byte accesses.  The element is read twice and written once.
namespace mnf {

namespace {
constexpr int kNhwcThreads = 256;
constexpr int kNhwcMaxC = 256;

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&f)[8]);
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float (&f)[8]) {
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
}  // namespace

// statistics: every summation runs in a FIXED order (per thread over its pixels, per CTA over its pixel slots, and the last CTA
// of an image over the CTAs' partial sums), so the result does not depend on scheduling: two runs give identical bits.
// scratch: partial [N][n_chunks][2C] floats, then stats [N][2C] (sum, sum of squares interleaved per channel), then one counter per image
template <typename T>
__global__ void __launch_bounds__(kNhwcThreads)
instance_norm_nhwc_stats_kernel(const T* __restrict__ x, float* __restrict__ partial, float* __restrict__ stats,
                                unsigned int* __restrict__ counters, const int hw, const int C, const int pix_per_cta) {
  __shared__ float part[kNhwcThreads][16 + 1];
  __shared__ bool last;
  const int n_oct = C >> 3, n_slots = kNhwcThreads / n_oct;
  const int tid = threadIdx.x, oct = tid % n_oct, slot = tid / n_oct;
  const int n = blockIdx.y, n_chunks = gridDim.x;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  if (slot < n_slots) {
    const T* base = x + ((size_t)n * hw) * C + oct * 8;
    // four pixels per step: all loads of a step are issued before the first addition (memory-level parallelism: the loop was
    // one L2 / HBM round trip per pixel); the additions keep the pixel order, so the sums are bit-identical to the plain loop
    int p = p0 + slot;
    for (; p + 3 * n_slots < p1; p += 4 * n_slots) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8<T>(base + (size_t)(p + u * n_slots) * C, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += v[u][i]; ss[i] += v[u][i] * v[u][i]; }
    }
    for (; p < p1; p += n_slots) {
      float v[8];
      load8<T>(base + (size_t)p * C, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += v[i]; ss[i] += v[i] * v[i]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { part[tid][2 * i] = s[i]; part[tid][2 * i + 1] = ss[i]; }
  __syncthreads();
  float* my_partial = partial + ((size_t)n * n_chunks + blockIdx.x) * 2 * C;
  for (int j = tid; j < 2 * C; j += kNhwcThreads) {          // j = channel * 2 + (0: sum, 1: sum of squares)
    const int o = (j >> 1) >> 3, e = ((j >> 1) & 7) * 2 + (j & 1);
    float a = 0.f;
    for (int sl = 0; sl < n_slots; ++sl) a += part[sl * n_oct + o][e];
    my_partial[j] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(&counters[n], 1u) == (unsigned)(n_chunks - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    const float* img = partial + (size_t)n * n_chunks * 2 * C;
    // (ncu / launch list: this tail was most of the kernel -- one thread walked the image's ~200 partial sums with ONE dependent
    // L2 load in flight, ~250 cycles each.  Sixteen loads per step now; the additions keep the chunk order: same bits.)
    for (int j = tid; j < 2 * C; j += kNhwcThreads) {
      float a = 0.f;
      int c = 0;
      for (; c + 16 <= n_chunks; c += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(img + (size_t)(c + u) * 2 * C + j);
#pragma unroll
        for (int u = 0; u < 16; ++u) a += v[u];
      }
      for (; c < n_chunks; ++c) a += __ldcg(img + (size_t)c * 2 * C + j);
      stats[(size_t)n * 2 * C + j] = a;
    }
    if (tid == 0) counters[n] = 0u;                            // ready for the next call (stream-ordered)
  }
}

template <typename T>
__global__ void __launch_bounds__(kNhwcThreads)
instance_norm_nhwc_apply_kernel(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                const float* __restrict__ stats, const int hw, const int C, const int mode, const float eps,
                                const int pix_per_cta) {
  __shared__ float mr[kNhwcMaxC * 2];        // (mean, rstd) per channel of this image
  const int n = blockIdx.y, tid = threadIdx.x;
  const float inv_hw = 1.0f / (float)hw;
  for (int c = tid; c < C; c += kNhwcThreads) {
    const float mean = stats[((size_t)n * C + c) * 2] * inv_hw;
    const float var = fmaxf(stats[((size_t)n * C + c) * 2 + 1] * inv_hw - mean * mean, 0.f);
    mr[2 * c] = mean;
    mr[2 * c + 1] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int n_oct = C >> 3;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  const size_t img = (size_t)n * hw * C;
  const int total = (p1 - p0) * n_oct;
#pragma unroll 4
  for (int i = tid; i < total; i += kNhwcThreads) {
    const int p = p0 + i / n_oct, oct = i - (i / n_oct) * n_oct;
    const size_t off = img + (size_t)p * C + oct * 8;
    float v[8], r[8], o[8];
    load8<T>(x + off, v);
    if (mode == 2) load8<T>(res + off, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = (v[j] - mr[2 * (oct * 8 + j)]) * mr[2 * (oct * 8 + j) + 1];
      if (mode >= 1) a = fmaxf(a, 0.f);
      if (mode == 2) a = fmaxf(a + r[j], 0.f);
      o[j] = a;
    }
    store8<T>(y + off, o);
  }
}

int launch_instance_norm_nhwc(const void* x, const void* res, void* y, int is_f16, float* scratch, int64_t scratch_floats, int N, int hw,
                              int C, int mode, float eps, cudaStream_t s) {
  if (N <= 0 || hw <= 0) return MNF_OK;
  if (C % 8 != 0 || C > kNhwcMaxC || C < 8) { set_error("instance_norm_nhwc: C = %d (needs a multiple of 8, <= %d)", C, kNhwcMaxC); return MNF_EUNSUPPORTED; }
  // enough CTAs for two waves of the 148 SMs on the small maps, at most 512 pixels per CTA on the large ones
  int pix_per_cta = (int)(((int64_t)hw * N + 591) / 592);
  pix_per_cta = pix_per_cta < 32 ? 32 : (pix_per_cta > 512 ? 512 : pix_per_cta);
  dim3 grid((unsigned)((hw + pix_per_cta - 1) / pix_per_cta), (unsigned)N);
  // scratch layout: [counters: 64 words, zero between calls][stats N x 2C][partials N x chunks x 2C]
  if (N > 64 || 64 + (int64_t)N * 2 * C * (1 + grid.x) > scratch_floats) { set_error("instance_norm_nhwc: %d images x %d channels x %u chunks exceed the context scratch", N, C, grid.x); return MNF_EUNSUPPORTED; }
  unsigned int* counters = reinterpret_cast<unsigned int*>(scratch);
  float* stats_scratch = scratch + 64;
  float* partial = stats_scratch + (size_t)N * 2 * C;
  if (is_f16) {
    instance_norm_nhwc_stats_kernel<__half><<<grid, kNhwcThreads, 0, s>>>(reinterpret_cast<const __half*>(x), partial, stats_scratch, counters, hw, C, pix_per_cta);
    instance_norm_nhwc_apply_kernel<__half><<<grid, kNhwcThreads, 0, s>>>(reinterpret_cast<const __half*>(x), reinterpret_cast<const __half*>(res),
                                                                         reinterpret_cast<__half*>(y), stats_scratch, hw, C, mode, eps, pix_per_cta);
  } else {
    instance_norm_nhwc_stats_kernel<float><<<grid, kNhwcThreads, 0, s>>>(reinterpret_cast<const float*>(x), partial, stats_scratch, counters, hw, C, pix_per_cta);
    instance_norm_nhwc_apply_kernel<float><<<grid, kNhwcThreads, 0, s>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(res),
                                                                        reinterpret_cast<float*>(y), stats_scratch, hw, C, mode, eps, pix_per_cta);
  }
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
