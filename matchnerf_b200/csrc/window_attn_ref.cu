// K-attn, fp32 CUDA-core variant: GMFlow split-window single-head attention.
//
// Replaces single_head_split_window_attention / single_head_full_attention
// (models/gmflow/transformer.py:46-105 / :8-16).  The Swin roll, the window partition and the shifted-window
// mask (:19-43) are index arithmetic here: nothing is rolled, permuted or materialised, and the L x L score
// matrix never leaves the SM (online softmax over 64-key tiles).
//
// Full-precision cross-check for the tcgen05 kernel and the path for shapes that kernel does not cover.
#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kC = 128;
constexpr int kTile = 64;
constexpr int kLd = kC + 4;
constexpr int kLdP = kTile + 4;

struct AttnSmem {
  float q[kTile][kLd];
  float k[kTile][kLd];
  float v[kTile][kLd];
  float p[kTile][kLdP];
  int qtok[kTile], qreg[kTile];
  int ktok[kTile], kreg[kTile];
};

struct WinGeom {
  int h, w, wh, ww, sh, sw, splits;
};

// local index i inside window (wy, wx) -> token id in the un-rolled [h*w] order and its shift-region label
__device__ __forceinline__ void window_token(const WinGeom& g, int wy, int wx, int i, int& tok, int& reg) {
  const int ly = i / g.ww, lx = i - ly * g.ww;
  const int ry = wy * g.wh + ly, rx = wx * g.ww + lx;          // coordinates in the rolled frame
  const int oy = (ry + g.sh) % g.h, ox = (rx + g.sw) % g.w;    // roll by (-sh, -sw): rolled[r] = orig[r + s]
  tok = oy * g.w + ox;
  // transformer.py:25-36: slices [0,-win) [-win,-shift) [-shift,end) along each axis of the rolled frame
  const int ay = (ry >= g.h - g.wh) + (ry >= g.h - g.sh);
  const int ax = (rx >= g.w - g.ww) + (rx >= g.w - g.sw);
  reg = (g.sh | g.sw) ? ay * 3 + ax : 0;
}

}  // namespace

__global__ void __launch_bounds__(256, 1)
window_attn_ref_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       float* __restrict__ out, const WinGeom g, const float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int Lw = g.wh * g.ww;
  const int win = blockIdx.y % (g.splits * g.splits), b = blockIdx.y / (g.splits * g.splits);
  const int wy = win / g.splits, wx = win - wy * g.splits;
  const int q0 = blockIdx.x * kTile;
  const size_t boff = (size_t)b * g.h * g.w * kC;

  if (tid < kTile) {
    int tok = 0, reg = 0;
    if (q0 + tid < Lw) window_token(g, wy, wx, q0 + tid, tok, reg);
    sm.qtok[tid] = q0 + tid < Lw ? tok : -1;
    sm.qreg[tid] = reg;
  }
  __syncthreads();
  for (int i = tid; i < kTile * (kC / 4); i += blockDim.x) {
    const int r = i / (kC / 4), c4 = i - r * (kC / 4);
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sm.qtok[r] >= 0) val = __ldg(reinterpret_cast<const float4*>(q + boff + (size_t)sm.qtok[r] * kC) + c4);
    *reinterpret_cast<float4*>(&sm.q[r][c4 * 4]) = val;
  }

  float m_run[4], l_run[4], o[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY;
    l_run[r] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) o[r][c] = 0.f;
  }

  for (int k0 = 0; k0 < Lw; k0 += kTile) {
    __syncthreads();  // previous tile fully consumed (also orders the Q stage on the first pass)
    if (tid < kTile) {
      int tok = 0, reg = 0;
      if (k0 + tid < Lw) window_token(g, wy, wx, k0 + tid, tok, reg);
      sm.ktok[tid] = k0 + tid < Lw ? tok : -1;
      sm.kreg[tid] = reg;
    }
    __syncthreads();
    for (int i = tid; i < kTile * (kC / 4); i += blockDim.x) {
      const int r = i / (kC / 4), c4 = i - r * (kC / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (sm.ktok[r] >= 0) {
        kv = __ldg(reinterpret_cast<const float4*>(k + boff + (size_t)sm.ktok[r] * kC) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(v + boff + (size_t)sm.ktok[r] * kC) + c4);
      }
      *reinterpret_cast<float4*>(&sm.k[r][c4 * 4]) = kv;
      *reinterpret_cast<float4*>(&sm.v[r][c4 * 4]) = vv;
    }
    __syncthreads();

    // scores: rows ty*4 + r, key columns tx + 16*cc
    float sc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) sc[r][cc] = 0.f;
    for (int c0 = 0; c0 < kC; c0 += 4) {
      float4 a[4], bb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&sm.q[ty * 4 + r][c0]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) bb[cc] = *reinterpret_cast<const float4*>(&sm.k[tx + 16 * cc][c0]);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          sc[r][cc] = fmaf(a[r].x, bb[cc].x, fmaf(a[r].y, bb[cc].y, fmaf(a[r].z, bb[cc].z, fmaf(a[r].w, bb[cc].w, sc[r][cc]))));
    }
    // scale, mask, online softmax (row statistics shared by the 16 threads with equal ty)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = ty * 4 + r;
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int col = tx + 16 * cc;
        float s = sc[r][cc] * scale;
        if (sm.kreg[col] != sm.qreg[row]) s += -100.0f;          // transformer.py:41, :90
        if (sm.ktok[col] < 0) s = -INFINITY;
        sc[r][cc] = s;
        mx = fmaxf(mx, s);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[r], mx);
      const float corr = expf(m_run[r] - m_new);
      float sum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float pe = expf(sc[r][cc] - m_new);
        sm.p[row][tx + 16 * cc] = pe;
        sum += pe;
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      l_run[r] = l_run[r] * corr + sum;
      m_run[r] = m_new;
#pragma unroll
      for (int c = 0; c < 8; ++c) o[r][c] *= corr;
    }
    __syncthreads();
    // O[rows ty*4+r][cols tx*8 .. +7] += P . V
    for (int j0 = 0; j0 < kTile; j0 += 4) {
      float4 pr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pr[r] = *reinterpret_cast<const float4*>(&sm.p[ty * 4 + r][j0]);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 v0 = *reinterpret_cast<const float4*>(&sm.v[j0 + jj][tx * 8]);
        const float4 v1 = *reinterpret_cast<const float4*>(&sm.v[j0 + jj][tx * 8 + 4]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float pv = jj == 0 ? pr[r].x : (jj == 1 ? pr[r].y : (jj == 2 ? pr[r].z : pr[r].w));
          o[r][0] = fmaf(pv, v0.x, o[r][0]); o[r][1] = fmaf(pv, v0.y, o[r][1]);
          o[r][2] = fmaf(pv, v0.z, o[r][2]); o[r][3] = fmaf(pv, v0.w, o[r][3]);
          o[r][4] = fmaf(pv, v1.x, o[r][4]); o[r][5] = fmaf(pv, v1.y, o[r][5]);
          o[r][6] = fmaf(pv, v1.z, o[r][6]); o[r][7] = fmaf(pv, v1.w, o[r][7]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = ty * 4 + r;
    if (sm.qtok[row] < 0) continue;
    const float inv = 1.f / l_run[r];
    float* dst = out + boff + (size_t)sm.qtok[row] * kC + tx * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(o[r][0] * inv, o[r][1] * inv, o[r][2] * inv, o[r][3] * inv);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[r][4] * inv, o[r][5] * inv, o[r][6] * inv, o[r][7] * inv);
  }
}

int launch_window_attn_ref(const float* q, const float* k, const float* v, float* out, int B, int h, int w, int C,
                           int num_splits, int with_shift, cudaStream_t s) {
  WinGeom g;
  g.h = h; g.w = w; g.splits = num_splits;
  g.wh = h / num_splits; g.ww = w / num_splits;
  g.sh = (with_shift && num_splits > 1) ? g.wh / 2 : 0;
  g.sw = (with_shift && num_splits > 1) ? g.ww / 2 : 0;
  const int Lw = g.wh * g.ww;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.cur();
  if (!configured) {
    MNF_CUDA_TRY(cudaFuncSetAttribute(window_attn_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AttnSmem)));
    configured = true;
  }
  dim3 grid((Lw + kTile - 1) / kTile, B * num_splits * num_splits);
  window_attn_ref_kernel<<<grid, 256, sizeof(AttnSmem), s>>>(q, k, v, out, g, 1.0f / sqrtf((float)C));
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf
