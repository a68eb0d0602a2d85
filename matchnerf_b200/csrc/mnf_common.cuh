// Shared declarations for the matchnerf_b200 CUDA kernels (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/matchnerf_b200.h"

namespace mnf {

constexpr int kViews = 3;
constexpr int kWidth = 128;       // decoder.net_width
constexpr int kDepth = 6;         // decoder.net_depth
constexpr int kSkip = 4;          // decoder.skip = [4]
constexpr int kL3D = 10;          // decoder.posenc.L_3D
constexpr int kEnc = 3 + 6 * kL3D;  // 63
constexpr int kCond = MNF_COND_DIM; // 22
constexpr int kCondPad = MNF_COND_PAD;
constexpr int kFeatCh = MNF_FEAT_CH;
constexpr int kG0 = 2;            // encoder.cos_n_group[0]
constexpr int kG1 = 8;            // encoder.cos_n_group[1]
constexpr int kMaxSamples = 256;

// error plumbing (api.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define MNF_CUDA_TRY(expr)                                  \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return ::mnf::cuda_fail(_e, #expr); \
  } while (0)

// Per-device launch state.  cudaFuncSetAttribute and the SM count belong to a DEVICE, not to the process: a library used from
// one process on several GPUs (nn.DataParallel, one context per device) must configure every device it launches on.
// The slot of the CURRENT device is returned; the ABI entry points make the context's device current (api.cu DeviceGuard).
constexpr int kMaxDevices = 64;
template <typename T>
struct PerDevice {
  T v[kMaxDevices] = {};
  T& cur() {
    int d = 0;
    cudaGetDevice(&d);
    return v[(d >= 0 && d < kMaxDevices) ? d : 0];
  }
};

// Camera block handed to kernels by value (constant bank): everything the per-ray geometry needs.
struct DevCams {
  float w2c[kViews][12];
  float K[kViews][9];
  float nf[kViews][2];
  int local_radius, local_dilation;   // encoder.feature_sample_local_radius / _dilation (0 / 1 in every shipped config; > 0: gather_local.cu)
  float c2w[12];             // target camera->world
  float Kinv[9];             // target inverse intrinsics
  float tnear, tfar;
  int W, H;
};

struct DevRays {
  const int64_t* ray_idx;
  int64_t first_ray;
  const float* jitter;
  int64_t n_rays;
  // explicit-tensor variants (the reference's unfused public methods); all NULL on the fused render path
  const float* points = nullptr;   // [R*S][3] world-space sample points  (query_cond_info, models/matchnerf.py:209)
  const float* ndc = nullptr;      // [R*S][3] view-0 NDC sample points   (CondNeRF.forward points_3D, cond_nerf.py:52)
  const float* dirs = nullptr;     // [R*S][3] per-sample view direction  (CondNeRF.forward ray_unit)
  // tile-list mode of the v3 gather (fix-up pass of gather_tc.cu): recompute the 16 x 8 pixel tiles listed in tile_list
  // (tile = band * tiles_x + tx, band counted from band0) of the contiguous range [first_ray, first_ray + n_rays)
  const int* tile_list = nullptr;
  const int* tile_count = nullptr;
  int tiles_x = 0, band0 = 0;
};

// ---- per-ray geometry ---------------------------------------------------------------------
// misc/camera.py:255-278 (legacy): pixel centre at integer coords; ray = c2w*[K^-1 (x,y,1), 1] - centre.
__device__ __forceinline__ void cast_ray(const DevCams& c, int64_t pix, float o[3], float d[3]) {
  const uint32_t pu = (uint32_t)pix, wu = (uint32_t)c.W;   // pixel ids fit 32 bits; 64-bit div/mod is ~10x dearer
  const uint32_t yi = pu / wu;
  const float x = (float)(pu - yi * wu);
  const float y = (float)yi;
  float cam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    cam[i] = __fadd_rn(__fadd_rn(__fmul_rn(x, c.Kinv[i * 3 + 0]), __fmul_rn(y, c.Kinv[i * 3 + 1])), c.Kinv[i * 3 + 2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = c.c2w[i * 4 + 3];
    float p = __fmul_rn(cam[0], c.c2w[i * 4 + 0]);
    p = __fadd_rn(p, __fmul_rn(cam[1], c.c2w[i * 4 + 1]));
    p = __fadd_rn(p, __fmul_rn(cam[2], c.c2w[i * 4 + 2]));
    p = __fadd_rn(p, t);
    o[i] = t;
    d[i] = __fsub_rn(p, t);
  }
}

// models/matchnerf.py:163-181 (legacy): t_i = near + (i + u)/(S-1) * (far - near)
__device__ __forceinline__ float sample_depth(const DevCams& c, int i, int S, float u) {
  return __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)i, u), (float)(S - 1)), __fsub_rn(c.tfar, c.tnear)), c.tnear);
}

// misc/camera.py:351-379: world point -> (u, v, z) normalised by (W-1, H-1, near..far) in source view v.
__device__ __forceinline__ void project_ndc(const DevCams& c, int v, const float p[3], float& u, float& vv, float& z) {
  const float* E = c.w2c[v];
  const float* K = c.K[v];
  float cam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fmul_rn(p[0], E[i * 4 + 0]);
    a = __fadd_rn(a, __fmul_rn(p[1], E[i * 4 + 1]));
    a = __fadd_rn(a, __fmul_rn(p[2], E[i * 4 + 2]));
    cam[i] = __fadd_rn(a, E[i * 4 + 3]);
  }
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fmul_rn(cam[0], K[i * 3 + 0]);
    a = __fadd_rn(a, __fmul_rn(cam[1], K[i * 3 + 1]));
    q[i] = __fadd_rn(a, __fmul_rn(cam[2], K[i * 3 + 2]));
  }
  u = __fdiv_rn(__fdiv_rn(q[0], q[2]), (float)(c.W - 1));
  vv = __fdiv_rn(__fdiv_rn(q[1], q[2]), (float)(c.H - 1));
  z = __fdiv_rn(__fsub_rn(q[2], c.nf[v][0]), __fsub_rn(c.nf[v][1], c.nf[v][0]));
}

// Same projection with reciprocal-multiply instead of IEEE division (a few ulp).  Used only where the result feeds
// fp16 operands (decoder positional encoding); the gather keeps the exact form because mask decisions hang on it.
__device__ __forceinline__ void project_ndc_fast(const DevCams& c, int v, const float p[3], float& u, float& vv, float& z) {
  const float* E = c.w2c[v];
  const float* K = c.K[v];
  float cam[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) cam[i] = fmaf(p[2], E[i * 4 + 2], fmaf(p[1], E[i * 4 + 1], fmaf(p[0], E[i * 4 + 0], E[i * 4 + 3])));
  float q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = fmaf(cam[2], K[i * 3 + 2], fmaf(cam[1], K[i * 3 + 1], cam[0] * K[i * 3 + 0]));
  const float rz = __frcp_rn(q[2]);
  u = q[0] * rz * __frcp_rn((float)(c.W - 1));
  vv = q[1] * rz * __frcp_rn((float)(c.H - 1));
  z = (q[2] - c.nf[v][0]) * __frcp_rn(c.nf[v][1] - c.nf[v][0]);
}

// grid_sample(align_corners=True, padding_mode='border') coordinate: g = uv*2-1; i = ((g+1)/2)*(n-1), clipped.
__device__ __forceinline__ float grid_unnormalize(float g, int n) {
  float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(n - 1));
  return fminf(fmaxf(i, 0.0f), (float)(n - 1));
}

// kernel launchers (defined in the .cu files; called from api.cu)
int gather_impl();   // 3 (default) or 4 (MNF_GATHER_IMPL=4, tensor-core experiment): selects the gather kernel AND the feature packing it reads
int launch_pack_features(const float* nchw, int V, int h, int w, __half* out, cudaStream_t s);
int launch_pack_images(const float* nchw, int V, int H, int W, float* out, cudaStream_t s);
// scratch: device ints owned by the context ([0] = counter, [1..] = tile list of the tensor-core gather's fix-up pass), or NULL
int launch_gather(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                  const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                  cudaStream_t s, int* scratch = nullptr, int scratch_ints = 0);

// encoder.feature_sample_local_radius > 0 (gather_local.cu): mean over the (2r+1)^2 dilated bilinear samples
int launch_gather_local(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1,
                        int w1, const float* images, float* cond_f32, __half* cond_f16, cudaStream_t s);

int launch_gather_bwd(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0, const __half* f1, int h1,
                      int w1, const float* dcond, float* g0, float* g1, cudaStream_t s);

struct DecoderWeightsF32;  // decoder_ref.cu
int launch_decoder_ref(const DevCams& cams, const DevRays& rays, const mnf_decoder_cfg& cfg,
                       const DecoderWeightsF32& w, const float* cond_f32, int setbg_opaque, float* out_rgb,
                       float* out_depth, float* out_opacity, float* aux, cudaStream_t s);

int launch_composite(const float* rgb, const float* sigma, const float* depth, int64_t n_rays, int S, int setbg_opaque,
                     float* out_rgb, float* out_depth, float* out_opacity, float* out_prob, cudaStream_t s);

int launch_instance_norm(const float* x, const float* res, float* y, int64_t planes, int hw, int mode, float eps, cudaStream_t s);
int launch_instance_norm_nhwc(const void* x, const void* res, void* y, int is_f16, float* scratch, int64_t scratch_floats, int N, int hw,
                              int C, int mode, float eps, cudaStream_t s);
int launch_token_layernorm(const void* x, int x_is_f16, const float* gamma, const float* beta, float eps, const float* residual,
                           const float* prefix, float* out_f32, __half* out_f16, int64_t rows, cudaStream_t s);

// token_block.cu: merge + LayerNorm (+ FFN + LayerNorm) + residual of a TransformerLayer, one tcgen05 kernel
int64_t token_block_weight_bytes(int with_ffn);
int launch_token_block_pack(const float* merge_w, const float* g1, const float* b1, const float* w1, const float* w2, const float* g2,
                            const float* b2, void* out, int with_ffn, cudaStream_t s);
int launch_token_block(const float* attn, const float* source, const void* weights, int with_ffn, float eps, float* out, int64_t T,
                       cudaStream_t s);

// metrics.cu: masked / cropped PSNR + SSIM sums of a rendered view (misc/metrics.py)
int launch_image_metrics(const float* pred, const float* gt, const unsigned char* mask, int H, int W, int y0, int x0, int rh, int rw,
                         float data_range, double* out4, cudaStream_t s);

int launch_window_attn_ref(const float* q, const float* k, const float* v, float* out, int B, int h, int w, int C,
                           int num_splits, int with_shift, cudaStream_t s);

}  // namespace mnf
