// K-mlp-composite, fp32 CUDA-core variant: conditional MLP + ray transformer + alpha compositing.
//
// Replaces CondNeRF.forward (models/rfdecoder/cond_nerf.py:52-100), MultiHeadAttention.forward
// (models/rfdecoder/ray_transformer.py:49-79), NeRF.composite (models/rfdecoder/nerf.py:101-124) and the
// view-0 NDC / direction prep of MatchNeRF.render (models/matchnerf.py:120-134).
//
// This is the full-precision kernel: every option of the shipped configs, any S <= 256, fp32 FMA math.
// It is the on-device cross-check for the tcgen05 kernel (decoder_tc.cu) and the path taken for
// configurations that kernel does not cover.  One CTA (256 threads) per ray; the trunk runs in chunks of
// 64 samples with activations in shared memory, then the ray transformer and the compositing scan run over
// the whole ray.
#include "decoder_weights.cuh"
#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kRows = 64;         // samples per trunk chunk
constexpr int kLdH = kWidth + 4;  // padded row stride of the activation tiles (keeps float4 alignment)
constexpr int kLdE = 64 + 4;

struct Smem {
  float h[kRows][kLdH];      // trunk activations / feature_linear output
  float g[kRows][kLdH];      // conditioning gate pts_bias(cond), later views_linears hidden
  float e[kRows][kLdE];      // positional encoding (63 + zero pad)
  float c[kRows][24];        // conditioning vector (22 + zero pad)
  float raw[kMaxSamples][16];   // alpha_linear output (ray transformer input)
  float kk[kMaxSamples][16];
  float vv[kMaxSamples][16];
  float rgb[kMaxSamples][3];
  float depth[kMaxSamples];
  float valid[kMaxSamples];  // number of views that see the sample
  float sigma[kMaxSamples];
  float scan[8];
  float dirvec[64];          // views_linears.0.weight[:,128:] . dir + bias  (per ray)
  float dir3[kRows][4];      // explicit per-sample view directions (CondNeRF.forward ray_unit), else unused
};

// acc[r][c] += sum_k in[(ty*4+r)][k] * Wt[k][tx*CPT + c]
template <int CPT>
__device__ __forceinline__ void gemm_rows4(const float* __restrict__ in, int ldi, const float* __restrict__ Wt,
                                           int K, float (&acc)[4][CPT], int ty, int tx) {
  constexpr int NOUT = CPT * 16;
  for (int k0 = 0; k0 < K; k0 += 4) {
    float4 a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(in + (ty * 4 + r) * ldi + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[CPT];
      const float* wr = Wt + (size_t)(k0 + kk) * NOUT + tx * CPT;
      if constexpr (CPT == 8) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr) + 1);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else if constexpr (CPT == 4) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
      } else {
        w[0] = __ldg(wr);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float av = kk == 0 ? a[r].x : (kk == 1 ? a[r].y : (kk == 2 ? a[r].z : a[r].w));
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
      }
    }
  }
}

__device__ __forceinline__ float act_fn(int kind, float x) { return kind == 0 ? fmaxf(x, 0.f) : (x > 0.f ? x : expm1f(x)); }

}  // namespace

__global__ void __launch_bounds__(256, 1)
decoder_ref_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const mnf_decoder_cfg cfg,
                   const DecoderWeightsF32 w, const float* __restrict__ cond, const int setbg_opaque,
                   float* __restrict__ out_rgb, float* __restrict__ out_depth, float* __restrict__ out_opacity,
                   float* __restrict__ aux) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const HeadParams& hp = *w.head;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int S = cfg.n_samples;
  const int64_t ray = blockIdx.x;
  const bool explicit_in = rays.ndc != nullptr;   // CondNeRF.forward on explicit tensors: NDC points + per-sample directions
  const int64_t pix = explicit_in ? 0 : (rays.ray_idx ? rays.ray_idx[ray] : rays.first_ray + ray);
  float o[3], d[3];
  cast_ray(cams, pix, o, d);

  // per-ray direction term of the colour head: normalize(ray) rotated into source view 0 (matchnerf.py:129-131)
  if (tid < 64) {
    const float nrm = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
    const float ux = d[0] / nrm, uy = d[1] / nrm, uz = d[2] / nrm;
    const float* E = cams.w2c[0];
    const float dx = ux * E[0] + uy * E[1] + uz * E[2];
    const float dy = ux * E[4] + uy * E[5] + uz * E[6];
    const float dz = ux * E[8] + uy * E[9] + uz * E[10];
    sm.dirvec[tid] = hp.views_dir[tid * 3 + 0] * dx + hp.views_dir[tid * 3 + 1] * dy + hp.views_dir[tid * 3 + 2] * dz + hp.views_b[tid];
  }

  for (int row0 = 0; row0 < S; row0 += kRows) {
    // ---- stage inputs of this chunk: positional encoding of the view-0 NDC point, conditioning vector
    if (tid < kRows) {
      const int s = row0 + tid;
      float x[3] = {0.f, 0.f, 0.f};
      if (s < S && explicit_in) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          x[i] = rays.ndc[((size_t)ray * S + s) * 3 + i];
          sm.dir3[tid][i] = rays.dirs[((size_t)ray * S + s) * 3 + i];
        }
        sm.depth[s] = 0.f;
      } else if (s < S) {
        const float u = rays.jitter ? rays.jitter[ray * S + s] : 0.f;
        const float t = sample_depth(cams, s, S, u);
        float p[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));
        project_ndc(cams, 0, p, x[0], x[1], x[2]);
        sm.depth[s] = t;
      }
      float* e = sm.e[tid];
#pragma unroll
      for (int i = 0; i < 3; ++i) e[i] = x[i];
      for (int k = 0; k < kL3D; ++k) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float a = x[i] * f;                 // cond_nerf.py:108-116: k-major, sin block then cos block
          e[3 + k * 3 + i] = s < S ? sinf(a) : 0.f;
          e[3 + 3 * kL3D + k * 3 + i] = s < S ? cosf(a) : 0.f;
        }
      }
      e[63] = 0.f;
    }
    for (int i = tid; i < kRows * 24; i += blockDim.x) {
      const int r = i / 24, c = i - r * 24;
      const int s = row0 + r;
      sm.c[r][c] = (s < S && c < kCond) ? cond[((size_t)ray * S + s) * kCond + c] : 0.f;
    }
    __syncthreads();
    if (tid < kRows && row0 + tid < S) sm.valid[row0 + tid] = sm.c[tid][19] + sm.c[tid][20] + sm.c[tid][21];

    // ---- gate = pts_bias(cond)   (cond_nerf.py:62)
    {
      float acc[4][8] = {};
      gemm_rows4<8>(&sm.c[0][0], 24, w.gate_wt, 24, acc, ty, tx);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) sm.g[ty * 4 + r][tx * 8 + c] = acc[r][c] + __ldg(w.gate_b + tx * 8 + c);
    }
    // ---- trunk: h = relu((W_i h + b_i) * gate); layer 5 consumes [enc, h]   (cond_nerf.py:61-67)
    for (int l = 0; l < kDepth; ++l) {
      float acc[4][8] = {};
      if (l == 0) {
        gemm_rows4<8>(&sm.e[0][0], kLdE, w.wt[0], 64, acc, ty, tx);
      } else if (l == kSkip + 1) {
        gemm_rows4<8>(&sm.e[0][0], kLdE, w.wt[l], 64, acc, ty, tx);
        gemm_rows4<8>(&sm.h[0][0], kLdH, w.wt[l] + 64 * kWidth, kWidth, acc, ty, tx);
      } else {
        gemm_rows4<8>(&sm.h[0][0], kLdH, w.wt[l], kWidth, acc, ty, tx);
      }
      __syncthreads();  // all reads of h done (and gate written, for l == 0)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int col = tx * 8 + c;
          sm.h[ty * 4 + r][col] = fmaxf((acc[r][c] + __ldg(w.b[l] + col)) * sm.g[ty * 4 + r][col], 0.f);
        }
      __syncthreads();
    }
    // ---- alpha head: raw = act(alpha_linear(h)) (+ sinusoid table)   (cond_nerf.py:75-77)
    {
      float acc[4][1] = {};
      gemm_rows4<1>(&sm.h[0][0], kLdH, w.alpha_wt, kWidth, acc, ty, tx);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int s = row0 + ty * 4 + r;
        if (s < S) {
          float v = act_fn(cfg.raytrans_act, acc[r][0] + hp.alpha_b[tx]);
          if (cfg.raytrans_posenc) {  // cond_nerf.py:118-127: pos / 10000^(2*(j//2)/16), sin on even j, cos on odd j
            const double ang = (double)s / pow(10000.0, 2.0 * (double)(tx / 2) / 16.0);
            v += (float)((tx & 1) ? cos(ang) : sin(ang));
          }
          sm.raw[s][tx] = v;
        }
      }
    }
    // ---- colour head: feature_linear -> [feature, dir] -> views_linears -> rgb   (cond_nerf.py:90-95)
    {
      float acc[4][8] = {};
      gemm_rows4<8>(&sm.h[0][0], kLdH, w.feat_wt, kWidth, acc, ty, tx);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) sm.g[ty * 4 + r][tx * 8 + c] = acc[r][c] + __ldg(w.feat_b + tx * 8 + c);
      __syncthreads();
      float acc2[4][4] = {};
      gemm_rows4<4>(&sm.g[0][0], kLdH, w.views_wt, kWidth, acc2, ty, tx);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = tx * 4 + c;
          float dterm = sm.dirvec[col];
          if (explicit_in) {
            const float* d3 = sm.dir3[ty * 4 + r];
            dterm = hp.views_dir[col * 3 + 0] * d3[0] + hp.views_dir[col * 3 + 1] * d3[1] + hp.views_dir[col * 3 + 2] * d3[2] + hp.views_b[col];
          }
          sm.e[ty * 4 + r][col] = fmaxf(acc2[r][c] + dterm, 0.f);
        }
      __syncthreads();
      if (tid < kRows * 3) {
        const int r = tid / 3, c = tid - r * 3;
        const int s = row0 + r;
        if (s < S) {
          float a = hp.rgb_b[c];
          for (int k = 0; k < 64; ++k) a = fmaf(sm.e[r][k], hp.rgb_w[c * 64 + k], a);
          sm.rgb[s][c] = 1.f / (1.f + expf(-a));
        }
      }
    }
    __syncthreads();
  }

  // ---- ray transformer over the S samples of this ray (ray_transformer.py:49-79)
  float q[16], x[16];
  const int s = tid;
  if (s < S) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = sm.raw[s][i];
#pragma unroll
    for (int oi = 0; oi < 16; ++oi) {
      float aq = 0.f, ak = 0.f, av = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        aq = fmaf(x[i], hp.att_q[oi * 16 + i], aq);
        ak = fmaf(x[i], hp.att_k[oi * 16 + i], ak);
        av = fmaf(x[i], hp.att_v[oi * 16 + i], av);
      }
      q[oi] = aq * 0.5f;  // q / temperature, temperature = sqrt(d_k) = 2
      sm.kk[s][oi] = ak;
      sm.vv[s][oi] = av;
    }
  }
  __syncthreads();
  float sig = 0.f;
  if (s < S) {
    const bool row_valid = sm.valid[s] > 1.f;   // cond_nerf.py:83: mask = (num_valid_obs > 1); masks whole query rows
    float attn[16];
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) {
      float mx = -INFINITY;
      for (int j = 0; j < S; ++j) {
        float sc = 0.f;
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) sc = fmaf(q[hd * 4 + dd], sm.kk[j][hd * 4 + dd], sc);
        mx = fmaxf(mx, row_valid ? sc : -1e9f);
      }
      float den = 0.f, o4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < S; ++j) {
        float sc = 0.f;
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) sc = fmaf(q[hd * 4 + dd], sm.kk[j][hd * 4 + dd], sc);
        const float pe = expf((row_valid ? sc : -1e9f) - mx);
        den += pe;
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) o4[dd] = fmaf(pe, sm.vv[j][hd * 4 + dd], o4[dd]);
      }
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) attn[hd * 4 + dd] = o4[dd] / den;
    }
    float y[16], mu = 0.f;
#pragma unroll
    for (int oi = 0; oi < 16; ++oi) {
      float a = x[oi];  // residual
#pragma unroll
      for (int i = 0; i < 16; ++i) a = fmaf(attn[i], hp.att_fc[oi * 16 + i], a);
      y[oi] = a;
      mu += a;
    }
    mu *= (1.f / 16.f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) var += (y[i] - mu) * (y[i] - mu);
    const float rstd = rsqrtf(var * (1.f / 16.f) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 16; ++i) y[i] = (y[i] - mu) * rstd * hp.ln_w[i] + hp.ln_b[i];
    float acc = hp.oa2_b;
#pragma unroll
    for (int oi = 0; oi < 16; ++oi) {
      float a = hp.oa0_b[oi];
#pragma unroll
      for (int i = 0; i < 16; ++i) a = fmaf(y[i], hp.oa0_w[oi * 16 + i], a);
      acc = fmaf(act_fn(cfg.raytrans_act, a), hp.oa2_w[oi], acc);
    }
    sig = fmaxf(acc, 0.f);
    if (cfg.density_maskfill && sm.valid[s] < 1.f) sig = 0.f;   // cond_nerf.py:86-87
    if (aux) {
      float* a4 = aux + ((size_t)ray * S + s) * 4;
      a4[0] = sm.rgb[s][0]; a4[1] = sm.rgb[s][1]; a4[2] = sm.rgb[s][2]; a4[3] = sig;
    }
  }

  // ---- alpha compositing (nerf.py:101-124, wo_render_interval): exclusive prefix sum of sigma along the ray
  float incl = sig;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += n;
  }
  if (lane == 31) sm.scan[wid] = incl;
  __syncthreads();
  float base = 0.f;
  for (int i = 0; i < wid; ++i) base += sm.scan[i];
  const float excl = base + incl - sig;
  const float wgt = s < S ? expf(-excl) * (1.f - expf(-sig)) : 0.f;
  float part[5];
  part[0] = s < S ? wgt * sm.rgb[s][0] : 0.f;
  part[1] = s < S ? wgt * sm.rgb[s][1] : 0.f;
  part[2] = s < S ? wgt * sm.rgb[s][2] : 0.f;
  part[3] = s < S ? wgt * sm.depth[s] : 0.f;
  part[4] = wgt;
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < 5; ++i) sm.kk[wid][i] = part[i];
  __syncthreads();
  if (tid == 0 && out_rgb) {
    float tot[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int wv = 0; wv < 8; ++wv)
      for (int i = 0; i < 5; ++i) tot[i] += sm.kk[wv][i];
    const float bg = setbg_opaque ? 1.f - tot[4] : 0.f;
    out_rgb[ray * 3 + 0] = tot[0] + bg;
    out_rgb[ray * 3 + 1] = tot[1] + bg;
    out_rgb[ray * 3 + 2] = tot[2] + bg;
    out_depth[ray] = tot[3];
    out_opacity[ray] = tot[4];
  }
}

int launch_decoder_ref(const DevCams& cams, const DevRays& rays, const mnf_decoder_cfg& cfg,
                       const DecoderWeightsF32& w, const float* cond_f32, int setbg_opaque, float* out_rgb,
                       float* out_depth, float* out_opacity, float* aux, cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.cur();
  const size_t smem = sizeof(Smem);
  if (!configured) {
    MNF_CUDA_TRY(cudaFuncSetAttribute(decoder_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  decoder_ref_kernel<<<(unsigned)rays.n_rays, 256, smem, s>>>(cams, rays, cfg, w, cond_f32, setbg_opaque, out_rgb,
                                                               out_depth, out_opacity, aux);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
