// Layout packing kernels: reference NCHW fp32 tensors -> the channels-last layouts the gather kernel reads.
//
// Feature map layout: a source view's 256 channels are two 128-channel halves, one per image pair the view takes part
// in (models/matchnerf.py:192-205).  A texel is stored as 256 fp16 = 512 B = 32 slots of 16 B.
//
// Packing v3 (default, gather v3): slot l (= lane l of the gathering warp, one 512 B load instruction per texel) holds
//   half0[4l .. 4l+3] and half1[4l .. 4l+3]
// i.e. packed position p = 8*l + e  <->  channel (e < 4 ? 0 : 128) + 4*l + (e & 3).
// Every lane therefore owns the same channel indices of both halves of every view, which makes all three pair products
// (v0h0.v1h0, v0h1.v2h0, v1h1.v2h1) lane-local; a fine-scale cosine group (16 channels) is a run of 4 lanes and a
// coarse group (64 channels) a run of 16 lanes.
//
// Packing v4 (gather v4, gather_mma.cu): x-pair interleaved blocks, [V][h][ceil(w/2)][256 channels][2 texels] fp16:
// one 32-bit word holds the same channel of texels (2i, 2i+1) of a row, i.e. one k-pair of an mma.sync B fragment;
// channels in natural order (half0 = words 0..127, half1 = words 128..255).  A missing odd texel (w odd) is zero.
#include "mnf_common.cuh"

namespace mnf {

__global__ void pack_features_kernel(const float* __restrict__ in, __half* __restrict__ out, int hw) {
  // grid: (ceil(hw/32), V); block 256 threads.  Tile = 32 pixels x 256 channels through shared memory.
  __shared__ float tile[kFeatCh][33];
  const int v = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  const float* src = in + (size_t)v * kFeatCh * hw;
  for (int c = ty; c < kFeatCh; c += 8) {
    const int p = p0 + tx;
    tile[c][tx] = p < hw ? src[(size_t)c * hw + p] : 0.f;
  }
  __syncthreads();
  // each thread emits 16 B (8 packed channels) for one (pixel, 16-byte slot) pair; 32 pixels x 32 slots = 1024 items
  for (int item = threadIdx.x; item < 32 * 32; item += blockDim.x) {
    const int pix = item >> 5, slot = item & 31;
    const int p = p0 + pix;
    if (p >= hw) continue;
    __align__(16) __half vals[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (j < 4 ? 0 : 128) + 4 * slot + (j & 3);
      vals[j] = __float2half_rn(tile[c][pix]);
    }
    *reinterpret_cast<uint4*>(out + ((size_t)v * hw + p) * kFeatCh + slot * 8) = *reinterpret_cast<const uint4*>(vals);
  }
}

// v4: tile = 32 consecutive pixels (row-major) x 256 channels; a thread writes one channel of one pixel (or of an
// x-pair when the width is even, so that pairs never straddle rows)
#ifdef MNF_EXPERIMENTS
__global__ void pack_features_v4_kernel(const float* __restrict__ in, __half* __restrict__ out, int h, int w, int wp) {
  __shared__ float tile[kFeatCh][33];
  const int hw = h * w;
  const int v = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = in + (size_t)v * kFeatCh * hw;
  for (int c = ty; c < kFeatCh; c += 8) {
    const int p = p0 + tx;
    tile[c][tx] = p < hw ? src[(size_t)c * hw + p] : 0.f;
  }
  __syncthreads();
  __half* dst = out + (size_t)v * h * wp * (2 * kFeatCh);
  if ((w & 1) == 0) {
    for (int item = threadIdx.x; item < 16 * kFeatCh; item += blockDim.x) {
      const int pp = item >> 8, c = item & (kFeatCh - 1);
      const int p = p0 + 2 * pp;
      if (p >= hw) continue;
      const int y = p / w, x = p - y * w;
      const __half2 val = __floats2half2_rn(tile[c][2 * pp], tile[c][2 * pp + 1]);
      *reinterpret_cast<__half2*>(dst + ((size_t)y * wp + (x >> 1)) * (2 * kFeatCh) + 2 * c) = val;
    }
  } else {
    for (int item = threadIdx.x; item < 32 * kFeatCh; item += blockDim.x) {
      const int pix = item >> 8, c = item & (kFeatCh - 1);
      const int p = p0 + pix;
      if (p >= hw) continue;
      const int y = p / w, x = p - y * w;
      dst[((size_t)y * wp + (x >> 1)) * (2 * kFeatCh) + 2 * c + (x & 1)] = __float2half_rn(tile[c][pix]);
    }
  }
}
#endif

__global__ void pack_images_kernel(const float* __restrict__ in, float4* __restrict__ out, int hw, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = i / hw, p = i - v * hw;
  const float* src = in + (size_t)v * 3 * hw + p;
  out[i] = make_float4(src[0], src[hw], src[2 * (size_t)hw], 0.f);
}

int launch_pack_features(const float* nchw, int V, int h, int w, __half* out, cudaStream_t s) {
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, V);
#ifdef MNF_EXPERIMENTS
  if (gather_impl() == 4) {
    const int wp = (w + 1) / 2;
    const size_t body = (size_t)V * h * wp * 2 * kFeatCh;
    // zero tail (one block row + 4 blocks: 8 x 2 texel windows touching the last row / column stay in bounds); with an odd width
    // the unpaired texels of every row are zero as well
    if (w & 1) MNF_CUDA_TRY(cudaMemsetAsync(out, 0, body * sizeof(__half), s));
    MNF_CUDA_TRY(cudaMemsetAsync(out + body, 0, (size_t)(wp + 4) * 2 * kFeatCh * sizeof(__half), s));
    pack_features_v4_kernel<<<grid, 256, 0, s>>>(nchw, out, h, w, wp);
    MNF_CUDA_TRY(cudaGetLastError());
    return MNF_OK;
  }
#endif
  pack_features_kernel<<<grid, 256, 0, s>>>(nchw, out, hw);
  MNF_CUDA_TRY(cudaGetLastError());
  // zero tail of (w + 1) texels: the zero-weight taps of samples on the last row / column stay in bounds (gather v3)
  MNF_CUDA_TRY(cudaMemsetAsync(out + (size_t)V * hw * kFeatCh, 0, (size_t)(w + 1) * kFeatCh * sizeof(__half), s));
  return MNF_OK;
}

int launch_pack_images(const float* nchw, int V, int H, int W, float* out, cudaStream_t s) {
  const int hw = H * W, total = V * hw;
  pack_images_kernel<<<(total + 255) / 256, 256, 0, s>>>(nchw, reinterpret_cast<float4*>(out), hw, total);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
