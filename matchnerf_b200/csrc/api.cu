// C ABI of libmatchnerf_b200.so (see include/matchnerf_b200.h).  Host-side glue only: argument checking,
// weight packing at load time, kernel dispatch.  No torch types, no allocation on the per-ray path.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "decoder_weights.cuh"
#include "mnf_common.cuh"

namespace mnf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return MNF_ECUDA;
}

// decoder_tc.cu
struct DecoderWeightsTC;
int decoder_tc_pack(const float* host_params, const ParamOffsets& off, DecoderWeightsTC** out);
void decoder_tc_free(DecoderWeightsTC* w);
int launch_decoder_tc(const DevCams& cams, const DevRays& rays, const mnf_decoder_cfg& cfg, const DecoderWeightsTC* w,
                      const HeadParams* head, const __half* cond_f16, int setbg_opaque, float* out_rgb, float* out_depth,
                      float* out_opacity, float* aux, cudaStream_t s);
bool decoder_tc_supports(const mnf_decoder_cfg& cfg);
// window_attn_tc.cu
bool window_attn_tc_supports(int B, int h, int w, int C, int num_splits);
int launch_window_attn_tc(const float* q, const float* k, const float* v, float* out, int B, int h, int w, int C,
                          int num_splits, int with_shift, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int64_t window_attn_tc_workspace_bytes(int B, int h, int w, int num_splits);
int64_t window_attn_proj_weight_bytes();
int launch_window_attn_pack_proj_weights(const float* wq, const float* wk, const float* wv, void* out, cudaStream_t s);
int launch_window_attn_proj_tc(const float* source, const float* target, const void* proj_weights, float* out, int B, int h, int w,
                               int num_splits, int with_shift, int target_roll, void* workspace, int64_t workspace_bytes, cudaStream_t s);

}  // namespace mnf

using namespace mnf;

constexpr int kGatherScratchInts = 1 << 16;
constexpr int64_t kInScratchFloats = 1 << 20;  // context scratch of the NHWC instance norm: counters, statistics, per-CTA partial sums (4 MB)

struct mnf_ctx {
  int device = 0;
  bool loaded = false;
  std::vector<void*> allocs;       // device allocations owned by the ctx
  DecoderWeightsF32 wf32{};
  HeadParams* head_dev = nullptr;
  DecoderWeightsTC* wtc = nullptr;
  float* in_stats = nullptr;       // device: [N][C][2] sum / sum-of-squares scratch of the NHWC instance norm (one in flight per ctx)
  int* gather_scratch = nullptr;   // device: [0] counter + tile list of the tensor-core gather's fix-up pass (one gather in flight per ctx)
};

namespace {

// Makes the context's device current for the duration of an ABI call and restores the caller's device afterwards: a
// model on cuda:1 may be driven by a thread whose current device is 0 (nn.DataParallel, torch.cuda.device scopes).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const mnf_ctx* ctx);
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

DeviceGuard::DeviceGuard(const mnf_ctx* ctx) {
  if (!ctx) return;
  if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
}

int dev_upload(mnf_ctx* ctx, const std::vector<float>& host, const float** out) {
  void* p = nullptr;
  MNF_CUDA_TRY(cudaMalloc(&p, host.size() * sizeof(float)));
  ctx->allocs.push_back(p);
  MNF_CUDA_TRY(cudaMemcpy(p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  *out = reinterpret_cast<const float*>(p);
  return MNF_OK;
}

// [N][K] row-major -> [Kpad][N] (transposed, zero padded rows)
std::vector<float> transpose_pad(const float* w, int N, int K, int Kpad) {
  std::vector<float> t((size_t)Kpad * N, 0.f);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) t[(size_t)k * N + n] = w[(size_t)n * K + k];
  return t;
}

int fill_cams(const mnf_scene* sc, DevCams* c) {
  if (!sc) { set_error("scene is NULL"); return MNF_EINVAL; }
  if (sc->n_views != kViews) { set_error("n_views = %d unsupported (only 3 source views)", sc->n_views); return MNF_EUNSUPPORTED; }
  if (sc->H < 2 || sc->W < 2) { set_error("bad image size %dx%d", sc->H, sc->W); return MNF_EINVAL; }
  memcpy(c->w2c, sc->src_w2c, sizeof(c->w2c));
  memcpy(c->K, sc->src_K, sizeof(c->K));
  memcpy(c->nf, sc->src_near_far, sizeof(c->nf));
  memcpy(c->c2w, sc->tgt_c2w, sizeof(c->c2w));
  memcpy(c->Kinv, sc->tgt_Kinv, sizeof(c->Kinv));
  c->tnear = sc->tgt_near_far[0];
  c->tfar = sc->tgt_near_far[1];
  c->W = sc->W;
  c->H = sc->H;
  if (sc->sample_local_radius < 0 || sc->sample_local_radius > 8 || (sc->sample_local_radius > 0 && sc->sample_local_dilation < 1)) {
    set_error("sample_local_radius = %d / sample_local_dilation = %d unsupported (radius 0..8, dilation >= 1)", sc->sample_local_radius, sc->sample_local_dilation);
    return MNF_EUNSUPPORTED;
  }
  c->local_radius = sc->sample_local_radius;
  c->local_dilation = sc->sample_local_radius > 0 ? sc->sample_local_dilation : 1;
  return MNF_OK;
}

int fill_rays(const mnf_scene* sc, const mnf_rays* r, DevRays* d) {
  if (!r) { set_error("rays is NULL"); return MNF_EINVAL; }
  if (r->n_rays < 0) { set_error("n_rays < 0"); return MNF_EINVAL; }
  if (!r->ray_idx && (r->first_ray < 0 || r->first_ray + r->n_rays > (int64_t)sc->H * sc->W)) {
    set_error("ray range [%lld, %lld) outside the %dx%d image", (long long)r->first_ray, (long long)(r->first_ray + r->n_rays), sc->H, sc->W);
    return MNF_EINVAL;
  }
  d->ray_idx = r->ray_idx;
  d->first_ray = r->first_ray;
  d->jitter = r->jitter;
  d->n_rays = r->n_rays;
  return MNF_OK;
}

int check_cfg(const mnf_decoder_cfg* cfg) {
  if (!cfg) { set_error("decoder cfg is NULL"); return MNF_EINVAL; }
  if (cfg->n_samples < 2 || cfg->n_samples > kMaxSamples) {
    set_error("n_samples = %d outside [2, %d]", cfg->n_samples, kMaxSamples);
    return MNF_EUNSUPPORTED;
  }
  if (cfg->raytrans_act != 0 && cfg->raytrans_act != 1) { set_error("raytrans_act must be 0 (ReLU) or 1 (ELU)"); return MNF_EUNSUPPORTED; }
  return MNF_OK;
}

}  // namespace

extern "C" {

int32_t mnf_abi_version(void) { return MNF_ABI_VERSION; }
const char* mnf_last_error(void) { return g_err; }
int64_t mnf_decoder_param_count(void) { return param_offsets().total; }

int32_t mnf_ctx_create(int32_t device, mnf_ctx** out) {
  if (!out) { set_error("mnf_ctx_create: out is NULL"); return MNF_EINVAL; }
  int n = 0;
  MNF_CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) { set_error("mnf_ctx_create: device %d not in [0, %d)", device, n); return MNF_EINVAL; }
  cudaDeviceProp prop;
  MNF_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("mnf_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    return MNF_EUNSUPPORTED;
  }
  mnf_ctx* c = new mnf_ctx();
  c->device = device;
  {
    DeviceGuard dev_guard(c);
    if (cudaMalloc(reinterpret_cast<void**>(&c->gather_scratch), kGatherScratchInts * sizeof(int)) != cudaSuccess) c->gather_scratch = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&c->in_stats), kInScratchFloats * sizeof(float)) != cudaSuccess) c->in_stats = nullptr;
    else cudaMemset(c->in_stats, 0, 64 * sizeof(float));      // the per-image CTA counters start (and are left) at zero
  }
  *out = c;
  return MNF_OK;
}

int32_t mnf_ctx_destroy(mnf_ctx* ctx) {
  if (!ctx) return MNF_OK;
  {
    DeviceGuard dev_guard(ctx);
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->wtc) decoder_tc_free(ctx->wtc);
    if (ctx->gather_scratch) cudaFree(ctx->gather_scratch);
    if (ctx->in_stats) cudaFree(ctx->in_stats);
  }
  delete ctx;
  return MNF_OK;
}

int32_t mnf_decoder_load_host(mnf_ctx* ctx, const float* P, int64_t n_floats) {
  if (!ctx || !P) { set_error("mnf_decoder_load_host: NULL argument"); return MNF_EINVAL; }
  const ParamOffsets off = param_offsets();
  if (n_floats != off.total) {
    set_error("mnf_decoder_load_host: got %lld floats, the decoder has %lld", (long long)n_floats, (long long)off.total);
    return MNF_EINVAL;
  }
  DeviceGuard dev_guard(ctx);
  for (void* p : ctx->allocs) cudaFree(p);
  ctx->allocs.clear();
  if (ctx->wtc) { decoder_tc_free(ctx->wtc); ctx->wtc = nullptr; }
  ctx->loaded = false;
  int rc;
  DecoderWeightsF32& w = ctx->wf32;
  for (int l = 0; l < kDepth; ++l) {
    std::vector<float> t;
    if (l == 0) {
      t = transpose_pad(P + off.pts_w[0], kWidth, kEnc, 64);
    } else if (l == kSkip + 1) {
      // [128][63 enc | 128 h] -> rows 0..63 enc (row 63 zero), rows 64..191 h
      const int K = kEnc + kWidth;
      t.assign((size_t)192 * kWidth, 0.f);
      for (int n = 0; n < kWidth; ++n) {
        for (int k = 0; k < kEnc; ++k) t[(size_t)k * kWidth + n] = P[off.pts_w[l] + (size_t)n * K + k];
        for (int k = 0; k < kWidth; ++k) t[(size_t)(64 + k) * kWidth + n] = P[off.pts_w[l] + (size_t)n * K + kEnc + k];
      }
    } else {
      t = transpose_pad(P + off.pts_w[l], kWidth, kWidth, kWidth);
    }
    if ((rc = dev_upload(ctx, t, &w.wt[l]))) return rc;
    std::vector<float> b(P + off.pts_b[l], P + off.pts_b[l] + kWidth);
    if ((rc = dev_upload(ctx, b, &w.b[l]))) return rc;
  }
  if ((rc = dev_upload(ctx, transpose_pad(P + off.gate_w, kWidth, kCond, 24), &w.gate_wt))) return rc;
  if ((rc = dev_upload(ctx, std::vector<float>(P + off.gate_b, P + off.gate_b + kWidth), &w.gate_b))) return rc;
  if ((rc = dev_upload(ctx, transpose_pad(P + off.alpha_w, 16, kWidth, kWidth), &w.alpha_wt))) return rc;
  if ((rc = dev_upload(ctx, transpose_pad(P + off.feat_w, kWidth, kWidth, kWidth), &w.feat_wt))) return rc;
  if ((rc = dev_upload(ctx, std::vector<float>(P + off.feat_b, P + off.feat_b + kWidth), &w.feat_b))) return rc;
  {
    // views_linears.0.weight [64][131] = [feature 128 | dir 3]
    std::vector<float> vf((size_t)kWidth * 64);
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < kWidth; ++k) vf[(size_t)k * 64 + n] = P[off.views_w + (size_t)n * (kWidth + 3) + k];
    if ((rc = dev_upload(ctx, vf, &w.views_wt))) return rc;
  }
  HeadParams hp;
  memcpy(hp.att_q, P + off.att_q, sizeof(hp.att_q));
  memcpy(hp.att_k, P + off.att_k, sizeof(hp.att_k));
  memcpy(hp.att_v, P + off.att_v, sizeof(hp.att_v));
  memcpy(hp.att_fc, P + off.att_fc, sizeof(hp.att_fc));
  memcpy(hp.ln_w, P + off.ln_w, sizeof(hp.ln_w));
  memcpy(hp.ln_b, P + off.ln_b, sizeof(hp.ln_b));
  memcpy(hp.oa0_w, P + off.oa0_w, sizeof(hp.oa0_w));
  memcpy(hp.oa0_b, P + off.oa0_b, sizeof(hp.oa0_b));
  memcpy(hp.oa2_w, P + off.oa2_w, sizeof(hp.oa2_w));
  hp.oa2_b = P[off.oa2_b];
  memcpy(hp.alpha_b, P + off.alpha_b, sizeof(hp.alpha_b));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 3; ++k) hp.views_dir[n * 3 + k] = P[off.views_w + (size_t)n * (kWidth + 3) + kWidth + k];
  memcpy(hp.views_b, P + off.views_b, sizeof(hp.views_b));
  memcpy(hp.rgb_w, P + off.rgb_w, sizeof(hp.rgb_w));
  memcpy(hp.rgb_b, P + off.rgb_b, sizeof(hp.rgb_b));
  void* hd = nullptr;
  MNF_CUDA_TRY(cudaMalloc(&hd, sizeof(HeadParams)));
  ctx->allocs.push_back(hd);
  MNF_CUDA_TRY(cudaMemcpy(hd, &hp, sizeof(HeadParams), cudaMemcpyHostToDevice));
  ctx->head_dev = reinterpret_cast<HeadParams*>(hd);
  w.head = ctx->head_dev;
  if ((rc = decoder_tc_pack(P, off, &ctx->wtc))) return rc;
  MNF_CUDA_TRY(cudaDeviceSynchronize());
  ctx->loaded = true;
  return MNF_OK;
}

int64_t mnf_packed_feature_halves(int32_t V, int32_t h, int32_t w) {
  if (V <= 0 || h <= 0 || w <= 0) return 0;
  // covers every packing: v3 = V*h*w + w + 1 texels; x-pair blocks = (V*h + 1) * 2*ceil(w/2) + 8 texels (blocks + tail)
  const int64_t w2 = 2 * (((int64_t)w + 1) / 2);
  return (((int64_t)V * h + 1) * w2 + 8) * kFeatCh;
}

int32_t mnf_pack_features(mnf_ctx* ctx, const float* feat_nchw, int32_t V, int32_t h, int32_t w, void* out_packed, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !feat_nchw || !out_packed || V <= 0 || h <= 0 || w <= 0) { set_error("mnf_pack_features: bad argument"); return MNF_EINVAL; }
  if (((uintptr_t)out_packed & 15) != 0) { set_error("mnf_pack_features: out must be 16-byte aligned"); return MNF_EINVAL; }
  return launch_pack_features(feat_nchw, V, h, w, reinterpret_cast<__half*>(out_packed), (cudaStream_t)stream);
}

int32_t mnf_pack_images(mnf_ctx* ctx, const float* images_nchw, int32_t V, int32_t H, int32_t W, void* out_packed, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !images_nchw || !out_packed || V <= 0 || H <= 0 || W <= 0) { set_error("mnf_pack_images: bad argument"); return MNF_EINVAL; }
  if (((uintptr_t)out_packed & 15) != 0) { set_error("mnf_pack_images: out must be 16-byte aligned"); return MNF_EINVAL; }
  return launch_pack_images(images_nchw, V, H, W, reinterpret_cast<float*>(out_packed), (cudaStream_t)stream);
}

int32_t mnf_gather_cossim_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, int32_t n_samples, float* cond_f32,
                              void* cond_f16, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  DevCams cams;
  DevRays dr{};
  int rc;
  if ((rc = fill_cams(scene, &cams))) return rc;
  if ((rc = fill_rays(scene, rays, &dr))) return rc;
  if (n_samples < 2 || n_samples > kMaxSamples) { set_error("n_samples = %d outside [2, %d]", n_samples, kMaxSamples); return MNF_EUNSUPPORTED; }
  if (!scene->feat0 || !scene->feat1 || !scene->images) { set_error("scene feature maps / images missing"); return MNF_EINVAL; }
  if (!cond_f32 && !cond_f16) { set_error("no output buffer"); return MNF_EINVAL; }
  return launch_gather(cams, dr, n_samples, reinterpret_cast<const __half*>(scene->feat0), scene->h0, scene->w0,
                       reinterpret_cast<const __half*>(scene->feat1), scene->h1, scene->w1,
                       reinterpret_cast<const float*>(scene->images), cond_f32, reinterpret_cast<__half*>(cond_f16),
                       (cudaStream_t)stream, ctx->gather_scratch, ctx->gather_scratch ? kGatherScratchInts : 0);
}

int32_t mnf_gather_cossim_bwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, int32_t n_samples, const float* dcond_f32,
                              float* grad_feat0_packed, float* grad_feat1_packed, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  DevCams cams;
  DevRays dr{};
  int rc;
  if ((rc = fill_cams(scene, &cams))) return rc;
  if ((rc = fill_rays(scene, rays, &dr))) return rc;
  if (n_samples < 2 || n_samples > kMaxSamples) { set_error("n_samples = %d outside [2, %d]", n_samples, kMaxSamples); return MNF_EUNSUPPORTED; }
  if (!scene->feat0 || !scene->feat1) { set_error("scene feature maps missing"); return MNF_EINVAL; }
  if (cams.local_radius > 0) { set_error("mnf_gather_cossim_bwd: sample_local_radius > 0 has no backward kernel"); return MNF_EUNSUPPORTED; }
  if (!dcond_f32 || !grad_feat0_packed || !grad_feat1_packed) { set_error("mnf_gather_cossim_bwd: NULL gradient buffer"); return MNF_EINVAL; }
  if ((((uintptr_t)grad_feat0_packed | (uintptr_t)grad_feat1_packed) & 15) != 0) { set_error("gradient maps must be 16-byte aligned"); return MNF_EINVAL; }
  return launch_gather_bwd(cams, dr, n_samples, reinterpret_cast<const __half*>(scene->feat0), scene->h0, scene->w0,
                           reinterpret_cast<const __half*>(scene->feat1), scene->h1, scene->w1, dcond_f32, grad_feat0_packed,
                           grad_feat1_packed, (cudaStream_t)stream);
}

int32_t mnf_decoder_composite_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, const mnf_decoder_cfg* cfg,
                                  const float* cond_f32, const void* cond_f16, int32_t setbg_opaque, float* out_rgb,
                                  float* out_depth, float* out_opacity, float* aux_rgb_sigma, int32_t impl, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  if (!ctx->loaded) { set_error("decoder weights not loaded (call mnf_decoder_load_host)"); return MNF_ESTATE; }
  DevCams cams;
  DevRays dr{};
  int rc;
  if ((rc = fill_cams(scene, &cams))) return rc;
  if ((rc = fill_rays(scene, rays, &dr))) return rc;
  if ((rc = check_cfg(cfg))) return rc;
  if (rays->n_rays == 0) return MNF_OK;
  if (!out_rgb || !out_depth || !out_opacity) { set_error("output buffer missing"); return MNF_EINVAL; }
  if (impl == 0) impl = (cond_f16 && decoder_tc_supports(*cfg)) ? 2 : 1;
  if (impl == 2) {
    if (!decoder_tc_supports(*cfg)) { set_error("tcgen05 decoder does not cover this configuration (S=%d)", cfg->n_samples); return MNF_EUNSUPPORTED; }
    if (!cond_f16) { set_error("tcgen05 decoder needs cond_f16"); return MNF_EINVAL; }
    return launch_decoder_tc(cams, dr, *cfg, ctx->wtc, ctx->head_dev, reinterpret_cast<const __half*>(cond_f16), setbg_opaque,
                             out_rgb, out_depth, out_opacity, aux_rgb_sigma, (cudaStream_t)stream);
  }
  if (impl != 1) { set_error("impl must be 0, 1 or 2"); return MNF_EINVAL; }
  if (!cond_f32) { set_error("fp32 decoder needs cond_f32"); return MNF_EINVAL; }
  return launch_decoder_ref(cams, dr, *cfg, ctx->wf32, cond_f32, setbg_opaque, out_rgb, out_depth, out_opacity, aux_rgb_sigma,
                            (cudaStream_t)stream);
}

int32_t mnf_query_cond_points_fwd(mnf_ctx* ctx, const mnf_scene* scene, const float* points_world, int64_t n_rays, int32_t n_samples,
                                  float* cond_f32, void* cond_f16, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  DevCams cams;
  int rc;
  if ((rc = fill_cams(scene, &cams))) return rc;
  if (n_rays < 0) { set_error("n_rays < 0"); return MNF_EINVAL; }
  if (n_samples < 1 || n_samples > kMaxSamples) { set_error("n_samples = %d outside [1, %d]", n_samples, kMaxSamples); return MNF_EUNSUPPORTED; }
  if (n_rays == 0) return MNF_OK;
  if (!points_world) { set_error("points_world is NULL"); return MNF_EINVAL; }
  if (!scene->feat0 || !scene->feat1 || !scene->images) { set_error("scene feature maps / images missing"); return MNF_EINVAL; }
  if (!cond_f32 && !cond_f16) { set_error("no output buffer"); return MNF_EINVAL; }
  DevRays dr{};
  dr.n_rays = n_rays;
  dr.points = points_world;
  return launch_gather(cams, dr, n_samples, reinterpret_cast<const __half*>(scene->feat0), scene->h0, scene->w0,
                       reinterpret_cast<const __half*>(scene->feat1), scene->h1, scene->w1,
                       reinterpret_cast<const float*>(scene->images), cond_f32, reinterpret_cast<__half*>(cond_f16),
                       (cudaStream_t)stream);
}

int32_t mnf_decoder_samples_fwd(mnf_ctx* ctx, const mnf_decoder_cfg* cfg, const float* pts_ndc, const float* ray_unit,
                                const float* cond_f32, int64_t n_rays, float* out_rgb_sigma, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  if (!ctx->loaded) { set_error("decoder weights not loaded (call mnf_decoder_load_host)"); return MNF_ESTATE; }
  int rc;
  if ((rc = check_cfg(cfg))) return rc;
  if (n_rays < 0) { set_error("n_rays < 0"); return MNF_EINVAL; }
  if (n_rays == 0) return MNF_OK;
  if (!pts_ndc || !ray_unit || !cond_f32 || !out_rgb_sigma) { set_error("mnf_decoder_samples_fwd: NULL tensor"); return MNF_EINVAL; }
  DevCams cams{};          // geometry comes from the explicit tensors; keep the unused camera block benign
  cams.W = cams.H = 2;
  DevRays dr{};
  dr.n_rays = n_rays;
  dr.ndc = pts_ndc;
  dr.dirs = ray_unit;
  return launch_decoder_ref(cams, dr, *cfg, ctx->wf32, cond_f32, 0, nullptr, nullptr, nullptr, out_rgb_sigma, (cudaStream_t)stream);
}

int32_t mnf_composite_fwd(mnf_ctx* ctx, const float* rgb, const float* sigma, const float* depth, int64_t n_rays,
                          int32_t n_samples, int32_t setbg_opaque, float* out_rgb, float* out_depth, float* out_opacity,
                          float* out_prob, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx) { set_error("ctx is NULL"); return MNF_EINVAL; }
  if (n_rays < 0 || n_samples < 1) { set_error("mnf_composite_fwd: bad shape R=%lld S=%d", (long long)n_rays, n_samples); return MNF_EINVAL; }
  if (n_rays == 0) return MNF_OK;
  if (!rgb || !sigma || !depth || !out_rgb || !out_depth || !out_opacity) { set_error("mnf_composite_fwd: NULL tensor"); return MNF_EINVAL; }
  return launch_composite(rgb, sigma, depth, n_rays, n_samples, setbg_opaque, out_rgb, out_depth, out_opacity, out_prob,
                          (cudaStream_t)stream);
}

int64_t mnf_render_workspace_bytes(int64_t n_rays, int32_t n_samples, int32_t impl) {
  if (n_rays < 0 || n_samples < 0) return 0;
  const int64_t n = n_rays * (int64_t)n_samples;
  if (impl == 0) {
    mnf_decoder_cfg cfg{};
    cfg.n_samples = n_samples;
    impl = decoder_tc_supports(cfg) ? 2 : 1;
  }
  // ONE conditioning region, in the layout the selected decoder kernel reads: fp32 [N][22] or fp16 [N][32]
  const int64_t bytes = impl == 1 ? n * kCond * 4 : n * kCondPad * 2;
  return ((bytes + 255) / 256) * 256;
}

int32_t mnf_render_rays_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, const mnf_decoder_cfg* cfg,
                            int32_t setbg_opaque, float* out_rgb, float* out_depth, float* out_opacity, void* workspace,
                            int64_t workspace_bytes, int32_t impl, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !rays || !cfg) { set_error("NULL argument"); return MNF_EINVAL; }
  int rc;
  if ((rc = check_cfg(cfg))) return rc;
  if (rays->n_rays == 0) return MNF_OK;
  if (impl == 0) impl = decoder_tc_supports(*cfg) ? 2 : 1;
  if (impl != 1 && impl != 2) { set_error("impl must be 0, 1 or 2"); return MNF_EINVAL; }
  const int64_t need = mnf_render_workspace_bytes(rays->n_rays, cfg->n_samples, impl);
  if (!workspace || workspace_bytes < need) {
    set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return MNF_ENOMEM;
  }
  if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return MNF_EINVAL; }
  float* cond32 = reinterpret_cast<float*>(workspace);
  void* cond16 = workspace;
  rc = mnf_gather_cossim_fwd(ctx, scene, rays, cfg->n_samples, impl == 1 ? cond32 : nullptr, impl == 2 ? cond16 : nullptr, stream);
  if (rc) return rc;
  return mnf_decoder_composite_fwd(ctx, scene, rays, cfg, impl == 1 ? cond32 : nullptr, impl == 2 ? cond16 : nullptr, setbg_opaque,
                                   out_rgb, out_depth, out_opacity, nullptr, impl, stream);
}

int32_t mnf_instance_norm_fwd(mnf_ctx* ctx, const float* x, const float* residual, float* y, int64_t n_planes, int32_t hw,
                              int32_t mode, float eps, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !x || !y) { set_error("mnf_instance_norm_fwd: NULL argument"); return MNF_EINVAL; }
  if (n_planes < 0 || hw <= 0) { set_error("mnf_instance_norm_fwd: bad shape planes=%lld hw=%d", (long long)n_planes, hw); return MNF_EINVAL; }
  if (mode < 0 || mode > 2 || (mode == 2 && !residual)) { set_error("mnf_instance_norm_fwd: mode must be 0, 1 or 2 (2 needs a residual)"); return MNF_EINVAL; }
  return launch_instance_norm(x, mode == 2 ? residual : nullptr, y, n_planes, hw, mode, eps, (cudaStream_t)stream);
}

int32_t mnf_instance_norm_nhwc_fwd(mnf_ctx* ctx, const void* x, const void* residual, void* y, int32_t is_f16, int32_t n_images,
                                   int32_t hw, int32_t channels, int32_t mode, float eps, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !x || !y) { set_error("mnf_instance_norm_nhwc_fwd: NULL argument"); return MNF_EINVAL; }
  if (n_images < 0 || hw <= 0 || channels <= 0) { set_error("mnf_instance_norm_nhwc_fwd: bad shape"); return MNF_EINVAL; }
  if (mode < 0 || mode > 2 || (mode == 2 && !residual)) { set_error("mnf_instance_norm_nhwc_fwd: mode must be 0, 1 or 2 (2 needs a residual)"); return MNF_EINVAL; }
  if (!ctx->in_stats) { set_error("mnf_instance_norm_nhwc_fwd: the context has no scratch buffer"); return MNF_ESTATE; }
  if ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)residual) & 15) != 0) { set_error("mnf_instance_norm_nhwc_fwd: pointers must be 16-byte aligned"); return MNF_EINVAL; }
  return launch_instance_norm_nhwc(x, mode == 2 ? residual : nullptr, y, is_f16, ctx->in_stats, kInScratchFloats, n_images, hw, channels, mode,
                                   eps, (cudaStream_t)stream);
}

int32_t mnf_token_layernorm_fwd(mnf_ctx* ctx, const void* x, int32_t x_is_f16, const float* gamma, const float* beta, float eps,
                                const float* residual, const float* prefix, float* out_f32, void* out_f16, int64_t n_tokens,
                                int32_t channels, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !x || !gamma || !beta) { set_error("mnf_token_layernorm_fwd: NULL argument"); return MNF_EINVAL; }
  if (channels != 128 || n_tokens < 0) { set_error("mnf_token_layernorm_fwd: %d channels / %lld tokens (the kernel is built for 128 channels)", channels, (long long)n_tokens); return MNF_EUNSUPPORTED; }
  if ((out_f32 != nullptr) == (out_f16 != nullptr)) { set_error("mnf_token_layernorm_fwd: exactly one of out_f32 / out_f16"); return MNF_EINVAL; }
  if ((out_f32 && prefix) || (out_f16 && residual)) { set_error("mnf_token_layernorm_fwd: residual goes with out_f32, prefix with out_f16"); return MNF_EINVAL; }
  if ((((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)residual | (uintptr_t)prefix | (uintptr_t)out_f32 | (uintptr_t)out_f16) & 15) != 0) {
    set_error("mnf_token_layernorm_fwd: pointers must be 16-byte aligned");
    return MNF_EINVAL;
  }
  return launch_token_layernorm(x, x_is_f16, gamma, beta, eps, residual, prefix, out_f32, reinterpret_cast<__half*>(out_f16), n_tokens,
                                (cudaStream_t)stream);
}

int32_t mnf_image_metrics_fwd(mnf_ctx* ctx, const float* pred_hwc, const float* gt_hwc, const uint8_t* mask_hw, int32_t H, int32_t W,
                              int32_t y0, int32_t x0, int32_t region_h, int32_t region_w, float data_range, double* out_sums4, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !pred_hwc || !gt_hwc || !out_sums4) { set_error("mnf_image_metrics_fwd: NULL argument"); return MNF_EINVAL; }
  if (H <= 0 || W <= 0 || y0 < 0 || x0 < 0 || region_h < 0 || region_w < 0 || y0 + region_h > H || x0 + region_w > W) {
    set_error("mnf_image_metrics_fwd: region [%d, %d) x [%d, %d) outside the %d x %d image", y0, y0 + region_h, x0, x0 + region_w, H, W);
    return MNF_EINVAL;
  }
  if (!(data_range > 0.f)) { set_error("mnf_image_metrics_fwd: data_range must be positive"); return MNF_EINVAL; }
  if (((uintptr_t)out_sums4 & 7) != 0) { set_error("mnf_image_metrics_fwd: out_sums4 must be 8-byte aligned"); return MNF_EINVAL; }
  return launch_image_metrics(pred_hwc, gt_hwc, mask_hw, H, W, y0, x0, region_h, region_w, data_range, out_sums4, (cudaStream_t)stream);
}

int64_t mnf_token_block_weight_bytes(int32_t with_ffn) { return token_block_weight_bytes(with_ffn); }

int32_t mnf_token_block_pack_weights(mnf_ctx* ctx, const float* merge_w, const float* norm1_w, const float* norm1_b, const float* mlp0_w,
                                     const float* mlp2_w, const float* norm2_w, const float* norm2_b, void* out_packed, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !merge_w || !norm1_w || !norm1_b || !out_packed) { set_error("mnf_token_block_pack_weights: NULL argument"); return MNF_EINVAL; }
  const int with_ffn = mlp0_w != nullptr;
  if (with_ffn && (!mlp2_w || !norm2_w || !norm2_b)) { set_error("mnf_token_block_pack_weights: mlp0_w given without mlp2_w / norm2"); return MNF_EINVAL; }
  if ((((uintptr_t)merge_w | (uintptr_t)mlp0_w | (uintptr_t)mlp2_w | (uintptr_t)out_packed) & 15) != 0) {
    set_error("mnf_token_block_pack_weights: pointers must be 16-byte aligned");
    return MNF_EINVAL;
  }
  return launch_token_block_pack(merge_w, norm1_w, norm1_b, mlp0_w, mlp2_w, norm2_w, norm2_b, out_packed, with_ffn, (cudaStream_t)stream);
}

int32_t mnf_token_block_fwd(mnf_ctx* ctx, const float* attn_out, const float* source, const void* weights_packed, int32_t with_ffn,
                            float eps, float* out, int64_t n_tokens, int32_t channels, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !attn_out || !source || !weights_packed || !out) { set_error("mnf_token_block_fwd: NULL argument"); return MNF_EINVAL; }
  if (channels != 128 || n_tokens < 0) { set_error("mnf_token_block_fwd: %d channels / %lld tokens (the kernel is built for d_model 128, ffn expansion 4)", channels, (long long)n_tokens); return MNF_EUNSUPPORTED; }
  if ((((uintptr_t)attn_out | (uintptr_t)source | (uintptr_t)weights_packed | (uintptr_t)out) & 15) != 0) {
    set_error("mnf_token_block_fwd: pointers must be 16-byte aligned");
    return MNF_EINVAL;
  }
  return launch_token_block(attn_out, source, weights_packed, with_ffn != 0, eps, out, n_tokens, (cudaStream_t)stream);
}

int64_t mnf_window_attn_workspace_bytes(int32_t B, int32_t h, int32_t w, int32_t num_splits) {
  return window_attn_tc_workspace_bytes(B, h, w, num_splits);
}

int64_t mnf_window_attn_proj_weight_bytes(void) { return window_attn_proj_weight_bytes(); }

int32_t mnf_window_attn_pack_proj_weights(mnf_ctx* ctx, const float* q_proj_w, const float* k_proj_w, const float* v_proj_w, void* out_packed,
                                          void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !q_proj_w || !k_proj_w || !v_proj_w || !out_packed) { set_error("mnf_window_attn_pack_proj_weights: NULL argument"); return MNF_EINVAL; }
  if ((((uintptr_t)q_proj_w | (uintptr_t)k_proj_w | (uintptr_t)v_proj_w | (uintptr_t)out_packed) & 15) != 0) {
    set_error("mnf_window_attn_pack_proj_weights: pointers must be 16-byte aligned");
    return MNF_EINVAL;
  }
  return launch_window_attn_pack_proj_weights(q_proj_w, k_proj_w, v_proj_w, out_packed, (cudaStream_t)stream);
}

int32_t mnf_window_attn_proj_fwd(mnf_ctx* ctx, const float* source, const float* target, const void* proj_weights_packed, float* out,
                                 int32_t B, int32_t h, int32_t w, int32_t C, int32_t num_splits, int32_t with_shift, int32_t target_batch_roll,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !source || !target || !proj_weights_packed || !out) { set_error("mnf_window_attn_proj_fwd: NULL argument"); return MNF_EINVAL; }
  if (C != 128) { set_error("mnf_window_attn_proj_fwd: C = %d unsupported (feature_channels is 128)", C); return MNF_EUNSUPPORTED; }
  if (B <= 0 || h <= 0 || w <= 0 || num_splits <= 0 || h % num_splits || w % num_splits) {
    set_error("mnf_window_attn_proj_fwd: bad shape B=%d h=%d w=%d splits=%d", B, h, w, num_splits);
    return MNF_EINVAL;
  }
  if ((((uintptr_t)source | (uintptr_t)target | (uintptr_t)proj_weights_packed | (uintptr_t)out) & 15) != 0) {
    set_error("mnf_window_attn_proj_fwd: pointers must be 16-byte aligned");
    return MNF_EINVAL;
  }
  return launch_window_attn_proj_tc(source, target, proj_weights_packed, out, B, h, w, num_splits, with_shift, target_batch_roll, workspace, workspace_bytes,
                                    (cudaStream_t)stream);
}

int32_t mnf_window_attn_fwd(mnf_ctx* ctx, const float* q, const float* k, const float* v, float* out, int32_t B, int32_t h,
                            int32_t w, int32_t C, int32_t num_splits, int32_t with_shift, int32_t impl, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  DeviceGuard dev_guard(ctx);
  if (!ctx || !q || !k || !v || !out) { set_error("mnf_window_attn_fwd: NULL argument"); return MNF_EINVAL; }
  if (C != 128) { set_error("mnf_window_attn_fwd: C = %d unsupported (feature_channels is 128)", C); return MNF_EUNSUPPORTED; }
  if (B <= 0 || h <= 0 || w <= 0 || num_splits <= 0 || h % num_splits || w % num_splits) {
    set_error("mnf_window_attn_fwd: bad shape B=%d h=%d w=%d splits=%d", B, h, w, num_splits);
    return MNF_EINVAL;
  }
  if (impl == 0) impl = window_attn_tc_supports(B, h, w, C, num_splits) ? 2 : 1;
  if (impl == 2) {
    if (!window_attn_tc_supports(B, h, w, C, num_splits)) { set_error("tcgen05 attention does not cover this shape"); return MNF_EUNSUPPORTED; }
    return launch_window_attn_tc(q, k, v, out, B, h, w, C, num_splits, with_shift, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  if (impl != 1) { set_error("impl must be 0, 1 or 2"); return MNF_EINVAL; }
  return launch_window_attn_ref(q, k, v, out, B, h, w, C, num_splits, with_shift, (cudaStream_t)stream);
}

}  // extern "C"
