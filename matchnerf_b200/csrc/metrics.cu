// Image metrics of the evaluation loop on the GPU: masked / cropped PSNR and SSIM of a rendered view against the ground truth.
//
// Replaces EvalTools.set_inputs / get_psnr / get_ssim (misc/metrics.py:19-46), which the reference evaluates on the host with numpy
// and scikit-image 0.19.2 (`structural_similarity(pred, gt, channel_axis=-1)`: 7 x 7 uniform window, sample covariance, K1 = 0.01,
// K2 = 0.03, data_range = 2 for float images -- skimage's dtype range (-1, 1) --, the (win - 1) / 2 = 3 pixel border cropped, mean
// over pixels and channels) after copying both images off the device.  Here the rendered image never leaves the GPU: one launch
// returns the four sums the two numbers are made of.
//
// The evaluated region is a rectangle of the image (the whole image, or the reference's centre crop to 80 %); with a mask
// (DTU: depth == 0) masked pixels count as 0 in BOTH images for SSIM and are left out of the PSNR mean (misc/metrics.py:23-41).
// A CTA covers a 16 x 16 pixel tile of the region: the tile + 3-pixel halo of both images goes to shared memory, a thread
// evaluates its pixel's 49-tap window sums (x, y, x^2, y^2, xy per channel) and the SSIM map value; fp32 window arithmetic, fp64
// accumulation across pixels (warp shuffles, shared memory, one double atomicAdd per CTA and quantity).
#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kMt = 16;                 // tile edge
constexpr int kWin = 7, kPad = 3;
constexpr int kHalo = kMt + 2 * kPad;   // 22

__global__ void __launch_bounds__(kMt * kMt)
image_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const unsigned char* __restrict__ mask, const int W,
                     const int y0, const int x0, const int rh, const int rw, const float c1, const float c2, double* __restrict__ out) {
  __shared__ float sp[3][kHalo][kHalo + 1], sg[3][kHalo][kHalo + 1];
  __shared__ double red[4][kMt * kMt / 32];
  const int tx = threadIdx.x & (kMt - 1), ty = threadIdx.x / kMt;
  const int bx = blockIdx.x * kMt, by = blockIdx.y * kMt;          // tile origin inside the region
  for (int i = threadIdx.x; i < kHalo * kHalo; i += kMt * kMt) {
    const int hy = i / kHalo, hx = i - hy * kHalo;
    const int ry = by + hy - kPad, rx = bx + hx - kPad;            // region coordinates
    float p[3] = {0.f, 0.f, 0.f}, g[3] = {0.f, 0.f, 0.f};
    if (ry >= 0 && ry < rh && rx >= 0 && rx < rw) {
      const size_t pix = (size_t)(y0 + ry) * W + (x0 + rx);
      if (!(mask && mask[pix])) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { p[c] = __ldg(pred + pix * 3 + c); g[c] = __ldg(gt + pix * 3 + c); }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { sp[c][hy][hx] = p[c]; sg[c][hy][hx] = g[c]; }
  }
  __syncthreads();
  const int ry = by + ty, rx = bx + tx;
  double v[4] = {0.0, 0.0, 0.0, 0.0};                               // squared error, its element count, SSIM map sum, its count
  if (ry < rh && rx < rw) {
    const size_t pix = (size_t)(y0 + ry) * W + (x0 + rx);
    if (!(mask && mask[pix])) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = sp[c][ty + kPad][tx + kPad] - sg[c][ty + kPad][tx + kPad];
        v[0] += (double)(d * d);
      }
      v[1] = 3.0;
    }
    if (ry >= kPad && ry < rh - kPad && rx >= kPad && rx < rw - kPad) {
      const float inv = 1.0f / (kWin * kWin), cov_norm = (float)(kWin * kWin) / (float)(kWin * kWin - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
        for (int dy = 0; dy < kWin; ++dy)
#pragma unroll
          for (int dx = 0; dx < kWin; ++dx) {
            const float a = sp[c][ty + dy][tx + dx], b = sg[c][ty + dy][tx + dx];
            sx += a; sy += b;
            sxx = fmaf(a, a, sxx); syy = fmaf(b, b, syy); sxy = fmaf(a, b, sxy);
          }
        const float ux = sx * inv, uy = sy * inv;
        const float vx = cov_norm * (sxx * inv - ux * ux), vy = cov_norm * (syy * inv - uy * uy), vxy = cov_norm * (sxy * inv - ux * uy);
        const float a1 = 2.f * ux * uy + c1, a2 = 2.f * vxy + c2, b1 = ux * ux + uy * uy + c1, b2 = vx + vy + c2;
        v[2] += (double)((a1 * a2) / (b1 * b2));
      }
      v[3] = 3.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < kMt * kMt / 32; ++w) a += red[threadIdx.x][w];
    if (a != 0.0) atomicAdd(out + threadIdx.x, a);
  }
}

}  // namespace

int launch_image_metrics(const float* pred, const float* gt, const unsigned char* mask, int H, int W, int y0, int x0, int rh, int rw,
                         float data_range, double* out4, cudaStream_t s) {
  MNF_CUDA_TRY(cudaMemsetAsync(out4, 0, 4 * sizeof(double), s));
  if (rh <= 0 || rw <= 0) return MNF_OK;
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  dim3 grid((unsigned)((rw + kMt - 1) / kMt), (unsigned)((rh + kMt - 1) / kMt));
  image_metrics_kernel<<<grid, kMt * kMt, 0, s>>>(pred, gt, mask, W, y0, x0, rh, rw, c1, c2, out4);
  MNF_CUDA_TRY(cudaGetLastError());
  (void)H;
  return MNF_OK;
}

}  // namespace mnf
