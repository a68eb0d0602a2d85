/*
 * matchnerf_b200.h -- C ABI of the B200-native MatchNeRF per-ray hot path.
 *
 * The reference (donydchen/matchnerf) is pure Python/PyTorch and has NO FFI of its own;
 * this header is the new seam placed directly under the Python methods that make up its hot
 * path.  Each entry point names the reference code it replaces.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless the name ends in _host.  The caller owns
 *     every buffer (inputs, outputs, workspaces); nothing on the per-ray path allocates,
 *     synchronises the host, or throws.  Work is enqueued on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream).
 *   - Every function returns 0 on success and a negative mnf_status on failure;
 *     mnf_last_error() returns a thread-local, human-readable reason.
 *   - A mnf_ctx is bound to one (process, device) pair and holds the packed decoder
 *     weights.  It is re-entrant across streams once loaded.
 *   - Shapes: V source views (must be 3), S samples per ray, R rays, N = R*S samples
 *     stored ray-major.  The decoder architecture is the one every shipped reference config
 *     uses (configs/base.yaml:29-39): width 128, depth 6, skip [4], L_3D 10, L_view 0,
 *     cos_n_group [2, 8]; other values are rejected with MNF_EUNSUPPORTED.
 */
#ifndef MATCHNERF_B200_H_
#define MATCHNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNF_ABI_VERSION 9

typedef enum mnf_status {
  MNF_OK = 0,
  MNF_EINVAL = -1,       /* bad argument (null pointer, size mismatch, misaligned buffer) */
  MNF_EUNSUPPORTED = -2, /* configuration outside the supported architecture */
  MNF_ECUDA = -3,        /* CUDA runtime error; see mnf_last_error() */
  MNF_ESTATE = -4,       /* e.g. decoder weights not loaded */
  MNF_ENOMEM = -5        /* caller workspace too small */
} mnf_status;

typedef struct mnf_ctx mnf_ctx;

/* Number of conditioning channels per sample: sum(cos_n_group)=10 + 3 views * RGB = 9 + 3 masks.
 * (models/rfdecoder/cond_nerf.py:18, :59) */
#define MNF_COND_DIM 22
/* cond rows are also produced as fp16 padded to 32 columns: the K extent of the gate MMA. */
#define MNF_COND_PAD 32
/* channels per source view per scale (two 128-channel halves, models/matchnerf.py:192-205) */
#define MNF_FEAT_CH 256

/* Decoder options that vary between the shipped configs (configs/*.yaml `decoder:` / `nerf:`). */
typedef struct mnf_decoder_cfg {
  int32_t n_samples;        /* nerf.sample_intvs (S); 2..256 */
  int32_t raytrans_act;     /* decoder.raytrans_act: 0 = ReLU, 1 = ELU */
  int32_t raytrans_posenc;  /* decoder.raytrans_posenc: add the sinusoid table to raw_alpha */
  int32_t density_maskfill; /* decoder.density_maskfill: sigma = 0 where no view sees the sample */
} mnf_decoder_cfg;

/* One encoded scene: 3 source views, their packed feature maps and cameras, plus the target camera.
 * Cameras are plain host values inside the struct (it is passed by pointer and read on the host). */
typedef struct mnf_scene {
  int32_t n_views;          /* must be 3 */
  int32_t H, W;             /* full-resolution size of source images and target view */
  int32_t h0, w0;           /* coarse feature map size (1/8 scale in the shipped configs) */
  int32_t h1, w1;           /* fine feature map size (1/4 scale) */
  const void* feat0;        /* device, packed by mnf_pack_features: [V][h0][w0][256] fp16 */
  const void* feat1;        /* device, packed by mnf_pack_features: [V][h1][w1][256] fp16 */
  const void* images;       /* device, packed by mnf_pack_images:   [V][H][W][4]   fp32 (RGB + pad) */
  float src_w2c[3][12];     /* world->camera [3x4] row-major per source view (batch.extrinsics[:, :3, :3, :]) */
  float src_K[3][9];        /* intrinsics per source view */
  float src_near_far[3][2];
  float tgt_c2w[12];        /* camera->world [3x4] of the target = float64 inverse of its w2c (misc/camera.py:231-240) */
  float tgt_Kinv[9];        /* inverse target intrinsics (misc/camera.py:221) */
  float tgt_near_far[2];
  /* encoder.feature_sample_local_radius / _dilation (models/gmflow/utils.py:136-162): 0 = plain bilinear samples (every shipped
   * config); r in 1..8 = each feature sample is the mean of the (2r+1)^2 bilinear samples at offsets {-r..r} * dilation around
   * it, positions scaled exactly as the reference does.  Forward only (mnf_gather_cossim_bwd rejects r > 0). */
  int32_t sample_local_radius;
  int32_t sample_local_dilation;
} mnf_scene;

/* Which rays to render: either an explicit pixel-id list or a contiguous row-major range. */
typedef struct mnf_rays {
  int64_t n_rays;           /* R */
  const int64_t* ray_idx;   /* device [R] pixel ids (y*W + x), or NULL to use first_ray + i */
  int64_t first_ray;        /* used when ray_idx == NULL (render_by_slices: models/matchnerf.py:153) */
  const float* jitter;      /* device [R][S] stratified offsets u in [0,1) (train mode), or NULL (models/matchnerf.py:168-172) */
} mnf_rays;

/* ---- library / context ------------------------------------------------------------------ */
int32_t mnf_abi_version(void);
const char* mnf_last_error(void);
int32_t mnf_ctx_create(int32_t device, mnf_ctx** out);
int32_t mnf_ctx_destroy(mnf_ctx* ctx);

/* Load the `nerf_dec` parameters (reference state_dict order, fp32, contiguous, HOST memory):
 *   pts_linears.{0..5}.{weight,bias}, pts_bias.{weight,bias}, views_linears.0.{weight,bias},
 *   alpha_linear.0.{weight,bias}, ray_attention.{w_qs,w_ks,w_vs,fc}.weight,
 *   ray_attention.layer_norm.{weight,bias}, out_alpha_linear.{0,2}.{weight,bias},
 *   feature_linear.{weight,bias}, rgb_linear.{weight,bias}            (130,324 floats)
 * Replaces: CondNeRF.define_network / load_state_dict (models/rfdecoder/cond_nerf.py:15-50).
 * Not a hot-path call: allocates device buffers inside the ctx and synchronises. */
int32_t mnf_decoder_load_host(mnf_ctx* ctx, const float* params_host, int64_t n_floats);
/* Number of floats mnf_decoder_load_host expects. */
int64_t mnf_decoder_param_count(void);

/* ---- layout packing (once per encoded scene) -------------------------------------------- */
/* fp32 NCHW feature maps [V][256][h][w] (the layout get_img_feat returns, models/matchnerf.py:183-207)
 * -> fp16 channels-last [V][h][w][256] with the two 128-channel halves interleaved in groups of 4
 * (see DESIGN.md "feature map layout"), followed by a zero tail of (w + 1) texels so that the
 * zero-weight bilinear taps of samples on the last row / column of a map stay in bounds.
 * out must hold mnf_packed_feature_halves(V, h, w) halves (slightly more than V*h*w*256), 16-byte aligned. */
int64_t mnf_packed_feature_halves(int32_t V, int32_t h, int32_t w);
int32_t mnf_pack_features(mnf_ctx* ctx, const float* feat_nchw, int32_t V, int32_t h, int32_t w,
                          void* out_packed, void* stream);
/* fp32 [V][3][H][W] images in [0,1] -> fp32 [V][H][W][4].  out must hold V*H*W*4 floats. */
int32_t mnf_pack_images(mnf_ctx* ctx, const float* images_nchw, int32_t V, int32_t H, int32_t W,
                        void* out_packed, void* stream);

/* ---- K-gather: epipolar projection + bilinear gather + grouped cosine similarity ---------- */
/* Replaces MatchNeRF.query_cond_info (models/matchnerf.py:209-293) incl. camera.get_coord_ref_ndc
 * (misc/camera.py:351-379), sample_depth (:163-181) and ray casting (misc/camera.py:255-286) for
 * the requested rays.  Writes, for each of the N = R*S samples, the 22 conditioning values in the
 * order CondNeRF.forward concatenates them (feat_info[10], color_info[9], mask_info[3]).
 *   cond_f32: [N][22] fp32 or NULL;  cond_f16: [N][32] fp16 (cols 22..31 zero) or NULL. */
int32_t mnf_gather_cossim_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, int32_t n_samples,
                              float* cond_f32, void* cond_f16, void* stream);
/* Backward of the above for the training step (coach.py:215-243: loss.backward() through query_cond_info): dcond_f32 [N][22] =
 * d(loss)/d(conditioning row) -- only columns 0..9 (the cosine similarities) reach parameters: F.grid_sample backward of
 * models/gmflow/utils.py:134 + nn.CosineSimilarity backward of models/matchnerf.py:268-271 -- is scattered (ADDED: the caller
 * zero-fills once per step) into fp32 gradient maps laid out like the packed feature maps, [V][h][w][256] channels-last in the
 * packed channel order of mnf_pack_features (position 8*l + e <-> channel (e < 4 ? 0 : 128) + 4*l + (e & 3)).  Same `scene` /
 * `rays` (incl. jitter) as the forward call. */
int32_t mnf_gather_cossim_bwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays, int32_t n_samples,
                              const float* dcond_f32, float* grad_feat0_packed, float* grad_feat1_packed, void* stream);

/* ---- K-mlp-composite: conditional MLP + ray transformer + alpha compositing -------------- */
/* Replaces CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100), MultiHeadAttention.forward
 * (models/rfdecoder/ray_transformer.py:49-79), NeRF.composite (models/rfdecoder/nerf.py:101-124) and
 * the view-0 NDC / ray-direction preparation in MatchNeRF.render (models/matchnerf.py:120-134).
 *   cond_f16 [N][32] fp16 as written by mnf_gather_cossim_fwd (used by the tensor-core kernel),
 *   cond_f32 [N][22] fp32 (used by the fp32 kernel); pass whichever the chosen impl needs (or both).
 *   out_rgb [R][3], out_depth [R], out_opacity [R] fp32.
 *   aux_rgb_sigma: optional [N][4] fp32 per-sample (r,g,b,sigma) for tests, or NULL.
 *   impl: 0 = auto (tensor-core kernel when the config allows), 1 = fp32 CUDA-core kernel,
 *         2 = tcgen05 kernel (error if unsupported for this config). */
int32_t mnf_decoder_composite_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays,
                                  const mnf_decoder_cfg* cfg, const float* cond_f32, const void* cond_f16,
                                  int32_t setbg_opaque, float* out_rgb, float* out_depth, float* out_opacity,
                                  float* aux_rgb_sigma, int32_t impl, void* stream);

/* ---- fused per-slice render: MatchNeRF.render (models/matchnerf.py:88-143) ---------------- */
/* Bytes of caller-provided workspace mnf_render_rays_fwd needs for R rays of S samples with decoder implementation `impl`
 * (0 = auto, 1 = fp32 kernel: [N][22] fp32 conditioning rows, 2 = tcgen05 kernel: [N][32] fp16 rows).  `cfg` may be NULL
 * (then only n_samples is used to resolve impl = 0). */
int64_t mnf_render_workspace_bytes(int64_t n_rays, int32_t n_samples, int32_t impl);
int32_t mnf_render_rays_fwd(mnf_ctx* ctx, const mnf_scene* scene, const mnf_rays* rays,
                            const mnf_decoder_cfg* cfg, int32_t setbg_opaque,
                            float* out_rgb, float* out_depth, float* out_opacity,
                            void* workspace, int64_t workspace_bytes, int32_t impl, void* stream);

/* ---- the reference's unfused per-sample methods on explicit tensors -------------------------- */
/* MatchNeRF.query_cond_info (models/matchnerf.py:209-293) on explicit world-space sample points
 * [R*S][3] (ray-major): same outputs as mnf_gather_cossim_fwd; the target camera of `scene` is unused. */
int32_t mnf_query_cond_points_fwd(mnf_ctx* ctx, const mnf_scene* scene, const float* points_world, int64_t n_rays,
                                  int32_t n_samples, float* cond_f32, void* cond_f16, void* stream);
/* CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100) on explicit tensors: pts_ndc [R*S][3] (points_3D),
 * ray_unit [R*S][3] (per-sample view direction), cond_f32 [R*S][22] (cat[feat_info, color_info, mask_info])
 * -> out_rgb_sigma [R*S][4] = (r, g, b, density) per sample.  fp32 CUDA-core kernel. */
int32_t mnf_decoder_samples_fwd(mnf_ctx* ctx, const mnf_decoder_cfg* cfg, const float* pts_ndc, const float* ray_unit,
                                const float* cond_f32, int64_t n_rays, float* out_rgb_sigma, void* stream);
/* NeRF.composite (models/rfdecoder/nerf.py:101-124, wo_render_interval): rgb [R][S][3], sigma [R][S], depth [R][S]
 * -> out_rgb [R][3], out_depth [R], out_opacity [R], out_prob [R][S] (or NULL). */
int32_t mnf_composite_fwd(mnf_ctx* ctx, const float* rgb, const float* sigma, const float* depth, int64_t n_rays,
                          int32_t n_samples, int32_t setbg_opaque, float* out_rgb, float* out_depth, float* out_opacity,
                          float* out_prob, void* stream);

/* ---- encoder helper: fused InstanceNorm2d + ReLU (+ residual) of the CNN backbone --------- */
/* x, y (and residual): fp32 [n_planes][hw] = contiguous NCHW with n_planes = N*C.  eps 1e-5, biased variance, no affine
 * (nn.InstanceNorm2d defaults, models/gmflow/backbone.py:13-25).
 *   mode 0: y = IN(x);  mode 1: y = relu(IN(x));  mode 2: y = relu(residual + relu(IN(x)))   (backbone.py:28-36) */
int32_t mnf_instance_norm_fwd(mnf_ctx* ctx, const float* x, const float* residual, float* y, int64_t n_planes,
                              int32_t hw, int32_t mode, float eps, void* stream);

/* Same three modes on channels-last activations: x, y, residual = fp32 or (is_f16) fp16 [n_images][hw][channels] (the memory of a
 * torch.channels_last NCHW tensor), channels a multiple of 8 (<= 256), n_images <= 64.  Used when the backbone runs
 * its cuDNN convolutions in NHWC (no layout-conversion kernels around them).  Two kernels (statistics in a fixed summation order: bit-reproducible;
 * normalise) over a context-owned scratch: one call in flight per context. */
int32_t mnf_instance_norm_nhwc_fwd(mnf_ctx* ctx, const void* x, const void* residual, void* y, int32_t is_f16, int32_t n_images,
                                   int32_t hw, int32_t channels, int32_t mode, float eps, void* stream);

/* ---- encoder helper: token LayerNorm of TransformerLayer.forward fused with what follows it -- */
/* x: [n_tokens][128], fp32 or (x_is_f16) fp16; gamma / beta: fp32 [128]; nn.LayerNorm semantics (biased variance, eps inside the
 * square root).  Exactly one output:
 *   out_f32 [n_tokens][128] = (residual ? residual : 0) + LN(x)      `source + message`, models/gmflow/transformer.py:176 and :183-185
 *   out_f16 [n_tokens][prefix ? 256 : 128] = half([prefix | LN(x)])  `cat([source, message])` as the FFN operand, :181
 * residual, prefix: fp32 [n_tokens][128] or NULL.  All pointers 16-byte aligned. */
int32_t mnf_token_layernorm_fwd(mnf_ctx* ctx, const void* x, int32_t x_is_f16, const float* gamma, const float* beta, float eps,
                                const float* residual, const float* prefix, float* out_f32, void* out_f16, int64_t n_tokens,
                                int32_t channels, void* stream);

/* ---- K-block: everything of TransformerLayer.forward after the attention, one tcgen05 kernel ---- */
/* Replaces models/gmflow/transformer.py:173-185 for d_model 128, ffn_dim_expansion 4:
 *     out = source + LN1(merge(attn_out))                                               (self-attention layer, no_ffn)
 *     out = source + LN2(mlp2(GELU(mlp0(cat[source, LN1(merge(attn_out))]))))             (cross-attention + FFN layer)
 * attn_out, source, out: [n_tokens][128] fp32 (out must not alias the inputs).  fp16 operands, fp32 accumulation; LayerNorm
 * (biased variance, eps inside the square root), erf-form GELU and the residual in fp32; the 1024-wide hidden tensor stays on chip.
 * The layer's parameters are packed once (fp16 operand tiles in streaming order + the LayerNorm vectors) by
 * mnf_token_block_pack_weights into mnf_token_block_weight_bytes(with_ffn) bytes of device memory, 16-byte aligned:
 *   merge_w [128][128], norm1_w / norm1_b [128]; with an FFN: mlp0_w [1024][256], mlp2_w [128][1024], norm2_w / norm2_b [128]
 *   (fp32, device, nn.Linear / nn.LayerNorm layouts); mlp0_w == NULL packs a no_ffn layer. */
int64_t mnf_token_block_weight_bytes(int32_t with_ffn);
int32_t mnf_token_block_pack_weights(mnf_ctx* ctx, const float* merge_w, const float* norm1_w, const float* norm1_b, const float* mlp0_w,
                                     const float* mlp2_w, const float* norm2_w, const float* norm2_b, void* out_packed, void* stream);
int32_t mnf_token_block_fwd(mnf_ctx* ctx, const float* attn_out, const float* source, const void* weights_packed, int32_t with_ffn,
                            float eps, float* out, int64_t n_tokens, int32_t channels, void* stream);

/* ---- K-attn: GMFlow split-window single-head attention ----------------------------------- */
/* Replaces single_head_split_window_attention / single_head_full_attention
 * (models/gmflow/transformer.py:46-105 / :8-16) including the roll, the window partition and the
 * shifted-window mask (:19-43), none of which are materialised.
 *   q, k, v, out: [B][h*w][128] fp32, token order row-major over (h, w).
 *   num_splits: windows per axis (1 = full attention); with_shift: Swin shift of half a window.
 *   impl: 0 = auto, 1 = fp32 CUDA-core kernel, 2 = tcgen05 kernel.
 *   workspace: device scratch of mnf_window_attn_workspace_bytes(B, h, w, num_splits) bytes, 16-byte aligned, or NULL.
 *     With it the tcgen05 path pre-packs every Q / K / V tile once per call (fp16, swizzled operand images) and feeds
 *     the attention kernel with TMA-unit bulk copies; without it every CTA gathers and converts its own tiles. */
int64_t mnf_window_attn_workspace_bytes(int32_t B, int32_t h, int32_t w, int32_t num_splits);
int32_t mnf_window_attn_fwd(mnf_ctx* ctx, const float* q, const float* k, const float* v, float* out,
                            int32_t B, int32_t h, int32_t w, int32_t C, int32_t num_splits, int32_t with_shift,
                            int32_t impl, void* workspace, int64_t workspace_bytes, void* stream);

/* The same attention with the layer's projections fused in front of it (TransformerLayer.forward, models/gmflow/transformer.py:158-171):
 *     out = attention(q_proj(source), k_proj(target), v_proj(target))
 * source, target, out: [B][h*w][128] fp32.  One kernel gathers each window tile from source / target, multiplies it by the fp16
 * projection weight on the tensor cores and writes the attention kernel's operand images directly (no fp32 q / k / v tensors, no
 * pre-pack pass); then the tcgen05 attention kernel runs.  fp16 operands, fp32 accumulation.
 *   proj_weights_packed: mnf_window_attn_proj_weight_bytes() bytes written by mnf_window_attn_pack_proj_weights from the three
 *     nn.Linear weights [128][128] fp32 (device), once per parameter state;
 *   target_batch_roll r: keys / values of batch item b are taken from target[(b + r) mod B] -- FeatureTransformer.forward pairs every
 *     view with "the other view" through torch.cat([x[b:], x[:b]]) (transformer.py:331); r = B/2 on the un-rolled tensor is the same;
 *   workspace: mnf_window_attn_workspace_bytes(B, h, w, num_splits) bytes, 16-byte aligned (required). */
int64_t mnf_window_attn_proj_weight_bytes(void);
int32_t mnf_window_attn_pack_proj_weights(mnf_ctx* ctx, const float* q_proj_w, const float* k_proj_w, const float* v_proj_w,
                                          void* out_packed, void* stream);
int32_t mnf_window_attn_proj_fwd(mnf_ctx* ctx, const float* source, const float* target, const void* proj_weights_packed, float* out,
                                 int32_t B, int32_t h, int32_t w, int32_t C, int32_t num_splits, int32_t with_shift, int32_t target_batch_roll,
                                 void* workspace, int64_t workspace_bytes, void* stream);

/* ---- evaluation metrics of a rendered view: PSNR and SSIM sums ------------------------------ */
/* Replaces EvalTools.set_inputs / get_psnr / get_ssim (misc/metrics.py:19-46; SSIM = scikit-image 0.19.2 structural_similarity with
 * its defaults: 7 x 7 uniform window, sample covariance, K1 0.01, K2 0.03, 3-pixel border cropped) without copying the images to the
 * host.  pred_hwc, gt_hwc: [H][W][3] fp32; mask_hw: [H][W] bytes (non-zero = masked OUT, e.g. DTU depth == 0) or NULL.
 * The metrics are evaluated on the region rows [y0, y0 + region_h) x columns [x0, x0 + region_w) (whole image, or the reference's
 * 80 % centre crop); masked pixels count as 0 in both images for SSIM and are left out of the PSNR mean.  data_range: the reference
 * gets 2.0 (skimage's range for float images).  out_sums4 (device, 4 doubles, overwritten):
 *   [0] sum of squared differences over unmasked region pixels x 3 channels   [1] their element count
 *   [2] sum of the SSIM map over interior pixels x 3 channels                 [3] its element count
 * PSNR = -10 log10([0] / [1]),  SSIM = [2] / [3]. */
int32_t mnf_image_metrics_fwd(mnf_ctx* ctx, const float* pred_hwc, const float* gt_hwc, const uint8_t* mask_hw, int32_t H, int32_t W,
                              int32_t y0, int32_t x0, int32_t region_h, int32_t region_w, float data_range, double* out_sums4,
                              void* stream);

/* ---- self tests of the tcgen05 building blocks (used by tests/ on the GPU box) ------------- */
/* D[128][N] = A[128][K] * B[N][K]^T with fp16 operands, fp32 accumulate, one CTA.
 * mode 0: A and B from shared memory (SWIZZLE_128B K-major descriptors);
 * mode 1: A staged in tensor memory (tcgen05.st), B from shared memory;
 * mode 2: as mode 1 with B given as [K][N] and consumed as an MN-major operand (N % 64 == 0).
 * a_f16 [128][K], b_f16 [N][K] ([K][N] in mode 2) device fp16; d_f32 [128][N] device fp32.
 * K % 64 == 0, K <= 256, N % 16 == 0, N <= 256. */
int32_t mnf_selftest_umma(const void* a_f16, const void* b_f16, float* d_f32, int32_t N, int32_t K,
                          int32_t mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MATCHNERF_B200_H_ */
