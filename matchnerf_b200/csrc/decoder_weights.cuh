// Device-resident decoder parameters, in the two packings the kernels consume.
#pragma once
#include "mnf_common.cuh"

namespace mnf {

// Offsets (in floats) of each reference parameter inside the flat host blob handed to
// mnf_decoder_load_host (reference state_dict order, models/rfdecoder/cond_nerf.py:15-50).
struct ParamOffsets {
  int64_t pts_w[kDepth], pts_b[kDepth];
  int64_t gate_w, gate_b;          // pts_bias
  int64_t views_w, views_b;        // views_linears.0  [64][131]
  int64_t alpha_w, alpha_b;        // alpha_linear.0   [16][128]
  int64_t att_q, att_k, att_v, att_fc, ln_w, ln_b;
  int64_t oa0_w, oa0_b, oa2_w, oa2_b;
  int64_t feat_w, feat_b;          // feature_linear   [128][128]
  int64_t rgb_w, rgb_b;            // rgb_linear       [3][64]
  int64_t total;
};

inline ParamOffsets param_offsets() {
  ParamOffsets o{};
  int64_t p = 0;
  auto take = [&](int64_t n) { int64_t r = p; p += n; return r; };
  for (int i = 0; i < kDepth; ++i) {
    const int k = i == 0 ? kEnc : (i == kSkip + 1 ? kWidth + kEnc : kWidth);
    o.pts_w[i] = take((int64_t)kWidth * k);
    o.pts_b[i] = take(kWidth);
  }
  o.gate_w = take((int64_t)kWidth * kCond);
  o.gate_b = take(kWidth);
  o.views_w = take(64 * (kWidth + 3));
  o.views_b = take(64);
  o.alpha_w = take(16 * kWidth);
  o.alpha_b = take(16);
  o.att_q = take(256);
  o.att_k = take(256);
  o.att_v = take(256);
  o.att_fc = take(256);
  o.ln_w = take(16);
  o.ln_b = take(16);
  o.oa0_w = take(256);
  o.oa0_b = take(16);
  o.oa2_w = take(16);
  o.oa2_b = take(1);
  o.feat_w = take((int64_t)kWidth * kWidth);
  o.feat_b = take(kWidth);
  o.rgb_w = take(3 * 64);
  o.rgb_b = take(3);
  o.total = p;
  return o;
}

// Small per-ray-transformer / head parameters, passed to kernels by pointer (read through the
// constant/L1 path; 1.3 KB).
struct HeadParams {
  float att_q[256], att_k[256], att_v[256], att_fc[256];   // [out][in]
  float ln_w[16], ln_b[16];
  float oa0_w[256], oa0_b[16], oa2_w[16], oa2_b;
  float alpha_b[16];
  float views_dir[64 * 3];   // views_linears.0.weight[:, 128:131]
  float views_b[64];
  float rgb_w[3 * 64], rgb_b[3];
};

// fp32 packing for the CUDA-core kernel: every matrix transposed to [K][N] (K padded to a multiple of 4).
struct DecoderWeightsF32 {
  const float* wt[kDepth];   // layer 0: [64][128] (row 63 zero); 1..4: [128][128]; 5: [192][128] = enc rows (64) then h rows
  const float* b[kDepth];
  const float* gate_wt;      // [24][128] (rows 22,23 zero)
  const float* gate_b;
  const float* alpha_wt;     // [128][16]
  const float* feat_wt;      // [128][128]
  const float* feat_b;
  const float* views_wt;     // [128][64]  (feature part of views_linears.0)
  const HeadParams* head;    // device
};

}  // namespace mnf
