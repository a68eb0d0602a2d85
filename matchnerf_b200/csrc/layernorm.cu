// Token LayerNorm of the GMFlow transformer layer fused with what follows it (models/gmflow/transformer.py:147-185):
//   mode 0   out_f32[r] = residual[r] + LN(x[r])                    `source + message`   (:176 self-attention layer, :185 after the FFN)
//   mode 1   out_f16[r] = half([prefix[r] | LN(x[r])])              `cat([source, message])` as the fp16 operand of the FFN (:181)
// C = 128 channels: one warp per token, a lane owns four consecutive channels (one 16-byte load), mean and (biased) variance by two
// shuffle reductions on register-resident values -- nn.LayerNorm semantics (eps inside the square root, affine).  HBM-bound:
// 0.5-1.5 KB per token in and out; replaces ATen's LayerNorm (48 us per call at 30,720 tokens on B200: a third of the time of the
// whole transformer outside the attention) + add / cat / dtype-conversion kernels.
#include <cuda_fp16.h>

#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kLnC = 128;
constexpr int kLnWarps = 8;

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 pack4(float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
token_layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                       const float* __restrict__ residual, const float* __restrict__ prefix, float* __restrict__ out_f32,
                       __half* __restrict__ out_f16, const int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4 v = load4<T>(x + row * kLnC + 4 * lane);
  float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s * (1.0f / kLnC);
  const float4 d = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
  float q = (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q * (1.0f / kLnC) + eps);
  const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * lane), b = *reinterpret_cast<const float4*>(beta + 4 * lane);
  float4 y = make_float4(fmaf(d.x * rstd, g.x, b.x), fmaf(d.y * rstd, g.y, b.y), fmaf(d.z * rstd, g.z, b.z), fmaf(d.w * rstd, g.w, b.w));
  if (out_f32) {
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + row * kLnC + 4 * lane);
      y = make_float4(r.x + y.x, r.y + y.y, r.z + y.z, r.w + y.w);
    }
    *reinterpret_cast<float4*>(out_f32 + row * kLnC + 4 * lane) = y;
  } else {
    __half* o = out_f16 + row * (prefix ? 2 * kLnC : kLnC);
    if (prefix) {
      *reinterpret_cast<uint2*>(o + 4 * lane) = pack4(*reinterpret_cast<const float4*>(prefix + row * kLnC + 4 * lane));
      o += kLnC;
    }
    *reinterpret_cast<uint2*>(o + 4 * lane) = pack4(y);
  }
}

}  // namespace

int launch_token_layernorm(const void* x, int x_is_f16, const float* gamma, const float* beta, float eps, const float* residual,
                           const float* prefix, float* out_f32, __half* out_f16, int64_t rows, cudaStream_t s) {
  if (rows <= 0) return MNF_OK;
  const unsigned grid = (unsigned)((rows + kLnWarps - 1) / kLnWarps);
  if (x_is_f16)
    token_layernorm_kernel<__half><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const __half*>(x), gamma, beta, eps, residual, prefix,
                                                                  out_f32, out_f16, rows);
  else
    token_layernorm_kernel<float><<<grid, kLnWarps * 32, 0, s>>>(reinterpret_cast<const float*>(x), gamma, beta, eps, residual, prefix,
                                                                 out_f32, out_f16, rows);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
