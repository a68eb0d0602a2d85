// K-attn, tcgen05 variant: GMFlow split-window single-head attention (head dim 128) as flash attention on the
// 5th-gen tensor cores.
//
// Replaces single_head_split_window_attention / single_head_full_attention
// (models/gmflow/transformer.py:46-105 / :8-16).  The Swin roll, the window partition and the 9-region shifted-window
// mask (:19-43) are index arithmetic on token ids; the L x L score matrix only ever exists as 128 x 128 tiles in
// tensor memory.
//
// One CTA = one 128-query tile of one window (2 CTAs per SM); it walks the window's keys in 128-key tiles:
//   S = Q K^T          tcgen05.mma  A = Q (smem, K-major)   B = K (smem, K-major)   D = S in tensor memory
//   softmax warps      thread = query row: online softmax on S (exp2 domain), P -> fp16 written over the dead S columns,
//                      running output O rescaled in tensor memory when a row maximum moved;
//   O += P V           tcgen05.mma  A = P (tensor memory)    B = V (smem, MN-major: rows = keys)
// Tensor memory: S / P columns [0,128), O columns [128,256).  Two kernels share this core: v3 gathers / converts its fp32
// operand rows itself (loader warps), v4 (default, needs the caller's workspace) is fed pre-packed operand images by
// cp.async.bulk.  (v1 / v2, the serial-phase versions, are described in DESIGN.md 4 and profiles/r01_ncu_summary.md.)
#include <cuda_fp16.h>

#include <cstdlib>

#include "mnf_common.cuh"
#include "tcgen05.cuh"

namespace mnf {

namespace {

constexpr int kC = 128;
constexpr int kTile = 128;
constexpr int kBlockBytes = kTile * 128;    // [128 rows][64 fp16]
constexpr int kColS = 0, kColO = 128;

struct WinGeomTc {
  int h, w, wh, ww, sh, sw, splits;
};

__device__ __forceinline__ void window_token(const WinGeomTc& g, int wy, int wx, int i, int& tok, int& reg) {
  const int ly = i / g.ww, lx = i - ly * g.ww;
  const int ry = wy * g.wh + ly, rx = wx * g.ww + lx;          // rolled frame
  const int oy = (ry + g.sh) % g.h, ox = (rx + g.sw) % g.w;    // original frame (roll by (-sh, -sw))
  tok = oy * g.w + ox;
  const int ay = (ry >= g.h - g.wh) + (ry >= g.h - g.sh);      // transformer.py:25-36
  const int ax = (rx >= g.w - g.ww) + (rx >= g.w - g.sw);
  reg = (g.sh | g.sw) ? ay * 3 + ax : 0;
}

__device__ __forceinline__ uint32_t pack_h2f(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// gather 128 rows x 128 fp32 channels (token ids in tok[], -1 = zero row) into two swizzled [128][64] fp16 blocks.
// A thread owns one 8-channel chunk (tid & 15) of rows (tid >> 4) + 16 i: all 16 of its 16-byte loads are issued before
// the first conversion, so the tile costs one L2 round trip instead of eight (v1 walked the rows one dependent load at
// a time and spent ~13k of its ~29k cycles per key tile waiting here).  256 threads.
__device__ __forceinline__ void load_rows_swizzled(const float* __restrict__ src, const int* tok, unsigned char* dst,
                                                   float scale, int tid) {
  const int ch = tid & 15, r0 = tid >> 4;
  int t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = tok[r0 + 16 * i];
  float4 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4* p = reinterpret_cast<const float4*>(src + (size_t)max(t[i], 0) * kC + ch * 8);
    a[i] = __ldg(p);
    b[i] = __ldg(p + 1);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (t[i] >= 0) {
      out.x = pack_h2f(a[i].x * scale, a[i].y * scale);
      out.y = pack_h2f(a[i].z * scale, a[i].w * scale);
      out.z = pack_h2f(b[i].x * scale, b[i].y * scale);
      out.w = pack_h2f(b[i].z * scale, b[i].w * scale);
    }
    *reinterpret_cast<uint4*>(dst + (ch >> 3) * kBlockBytes + tc::sw128_offset(r0 + 16 * i, (ch & 7) * 8)) = out;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// v3: warp-specialised pipeline.  ncu of v2: issue 20 %, 4.2 barrier + 5.3 long-scoreboard stalls per issue -- every key
// tile ran load -> QK -> softmax -> PV strictly one after the other, with half the warps idle in each phase.  Here
//   warps 0-3  softmax: thread = query row, as before;
//   warps 4-7  loaders: gather K of tile t+1 as soon as QK(t) has consumed the K buffer (i.e. underneath softmax(t)),
//              then V of tile t+1 as soon as PV(t) has consumed the V buffer; they also own the token / shift-region
//              bookkeeping (double-buffered bit masks);
//   warp 8     one elected lane issues the MMAs and commits them to mbarriers.
// Single K / V buffers (2 CTAs per SM still fit), five mbarriers, no __syncthreads inside the loop.
constexpr int kPipeThreads = 288;

struct AttnPipeSmem {
  alignas(1024) unsigned char q[2][kBlockBytes];
  unsigned char k[2][kBlockBytes];
  unsigned char v[2][kBlockBytes];
  int qtok[kTile];
  int ktok[kTile];
  uint32_t kmask[2][9][4];         // [tile parity][query region][word]: key lies in ANOTHER shift region
  uint32_t kinval[2][4];           // [tile parity][word]: key past the end of the window
  alignas(8) uint64_t bar_s;       // QK(t) complete            (tcgen05.commit)
  uint64_t bar_o;                  // PV(t) complete            (tcgen05.commit)
  uint64_t k_full;                 // K(t) in shared memory     (4 loader warps)
  uint64_t v_full;                 // V(t) in shared memory     (4 loader warps)
  uint64_t p_ready;                // P(t) in tensor memory, O rescaled   (4 softmax warps)
  uint32_t tmem_base;
};

// rows r0 + 8 i (i < 16) of a 128-row tile, one 8-channel chunk per thread, for a group of 128 loader threads; two batches
// of 8 rows so that at most 16 16-byte loads (64 registers) are in flight per thread
__device__ __forceinline__ void load_rows_swizzled_128(const float* __restrict__ src, const int* tok, unsigned char* dst, int ltid) {
  const int ch = ltid & 15, r0 = ltid >> 4;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    int t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = tok[r0 + 8 * (half * 8 + i)];
    float4 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4* p = reinterpret_cast<const float4*>(src + (size_t)max(t[i], 0) * kC + ch * 8);
      a[i] = __ldg(p);
      b[i] = __ldg(p + 1);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (t[i] >= 0) {
        out.x = pack_h2f(a[i].x, a[i].y);
        out.y = pack_h2f(a[i].z, a[i].w);
        out.z = pack_h2f(b[i].x, b[i].y);
        out.w = pack_h2f(b[i].z, b[i].w);
      }
      *reinterpret_cast<uint4*>(dst + (ch >> 3) * kBlockBytes + tc::sw128_offset(r0 + 8 * (half * 8 + i), (ch & 7) * 8)) = out;
    }
  }
}

__global__ void __launch_bounds__(kPipeThreads, 2)
window_attn_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      float* __restrict__ out, const WinGeomTc g) {
  extern __shared__ unsigned char smem_dyn[];
  AttnPipeSmem& sm = *reinterpret_cast<AttnPipeSmem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lw = g.wh * g.ww;
  const int win = blockIdx.y % (g.splits * g.splits), b = blockIdx.y / (g.splits * g.splits);
  const int wy = win / g.splits, wx = win - wy * g.splits;
  const int q0 = blockIdx.x * kTile;
  const size_t boff = (size_t)b * g.h * g.w * kC;
  const bool shifted = (g.sh | g.sw) != 0;
  const float kLog2e = 1.4426950408889634f;
  const int n_kt = (Lw + kTile - 1) / kTile;

  if (tid == 0) {
    tc::mbar_init(&sm.bar_s, 1);
    tc::mbar_init(&sm.bar_o, 1);
    tc::mbar_init(&sm.k_full, 4);
    tc::mbar_init(&sm.v_full, 4);
    tc::mbar_init(&sm.p_ready, 4);
    tc::fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc<256>(&sm.tmem_base);
  int my_qreg = 0;
  if (tid < kTile) {
    int tok = -1, reg = 0;
    if (q0 + tid < Lw) window_token(g, wy, wx, q0 + tid, tok, reg);
    sm.qtok[tid] = tok;
    my_qreg = reg;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  // Q, pre-scaled so that scores come out in the exp2 domain: (q . k) / sqrt(C) * log2(e)
  if (tid < 256) load_rows_swizzled(q + boff, sm.qtok, &sm.q[0][0], rsqrtf((float)kC) * kLog2e, tid);
  tc::fence_proxy_async_smem();
  __syncthreads();

  if (warp >= 4 && warp < 8) {
    // ================================================================== loaders
    const int ltid = tid - 128, lw = warp - 4;
    for (int t = 0; t < n_kt; ++t) {
      const int par = t & 1;
      // K buffer, token list and mask buffers of parity `par` are free once QK(t-1) is complete: that MMA was issued
      // after PV(t-2), which waited for softmax(t-2), the last reader of kmask[par]
      if (t > 0) tc::mbar_wait(&sm.bar_s, (t - 1) & 1);
      int tok = -1, reg = 0;
      if (t * kTile + ltid < Lw) window_token(g, wy, wx, t * kTile + ltid, tok, reg);
      sm.ktok[ltid] = tok;
      if (shifted) {
#pragma unroll
        for (int r = 0; r < 9; ++r) {
          const uint32_t other = __ballot_sync(0xffffffffu, reg != r);
          if (lane == 0) sm.kmask[par][r][lw] = other;
        }
      }
      const uint32_t inval = __ballot_sync(0xffffffffu, tok < 0);
      if (lane == 0) sm.kinval[par][lw] = inval;
      asm volatile("bar.sync 1, 128;" ::: "memory");             // token list complete (loader warps only)
      load_rows_swizzled_128(k + boff, sm.ktok, &sm.k[0][0], ltid);
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sm.k_full);
      if (t > 0) tc::mbar_wait(&sm.bar_o, (t - 1) & 1);          // V buffer free: PV(t-1) complete
      load_rows_swizzled_128(v + boff, sm.ktok, &sm.v[0][0], ltid);
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sm.v_full);
      asm volatile("bar.sync 1, 128;" ::: "memory");             // everyone is done with ktok before the next tile rewrites it
    }
  } else if (warp == 8) {
    // ================================================================== MMA issuer
    const uint32_t idesc_qk = tc::umma_idesc_f16(128, 128, 0), idesc_pv = tc::umma_idesc_f16(128, 128, 1);
    for (int kt = 0; kt < n_kt; ++kt) {
      tc::mbar_wait(&sm.k_full, kt & 1);
      if (kt > 0) tc::mbar_wait(&sm.bar_o, (kt - 1) & 1);        // S / P columns free, O accumulation ordered
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(&sm.q[ks >> 2][0]) + (ks & 3) * 32);
          const uint64_t bd = tc::umma_desc_sw128(tc::smem_u32(&sm.k[ks >> 2][0]) + (ks & 3) * 32);
          tc::umma_ss(tmem + kColS, ad, bd, idesc_qk, ks > 0);
        }
        tc::umma_commit(&sm.bar_s);
      }
      __syncwarp();
      tc::mbar_wait(&sm.p_ready, kt & 1);
      tc::mbar_wait(&sm.v_full, kt & 1);
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 keys per step: A = P columns [8 ks, 8 ks + 8), B = V rows [16 ks, 16 ks + 16)
          const uint64_t bd = tc::umma_desc_sw128_mn(tc::smem_u32(&sm.v[0][0]) + ks * 2048, kBlockBytes);
          tc::umma_ts(tmem + kColO, tmem + kColS + ks * 8, bd, idesc_pv, (kt | ks) ? 1u : 0u);
        }
        tc::umma_commit(&sm.bar_o);
      }
      __syncwarp();
    }
  } else {
    // ================================================================== softmax: thread = query row
    const int row = warp * 32 + lane;
    const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16);
    const float kMaskAdd = -100.0f * kLog2e;                     // transformer.py:41, :90
    float m_run = -INFINITY, l_run = 0.f;
    for (int kt = 0; kt < n_kt; ++kt) {
      const int par = kt & 1;
      const bool partial = Lw - kt * kTile < kTile;              // warp-uniform
      tc::mbar_wait(&sm.bar_s, par);
      tc::tc_fence_after_sync();
      // sweep 1: row maximum of the masked scores
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < kTile; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tb + kColS + c0, r);
        const uint32_t mk = shifted ? sm.kmask[par][my_qreg][c0 >> 5] : 0u;
        const uint32_t iv = partial ? sm.kinval[par][c0 >> 5] : 0u;
        tc::tmem_wait_ld();
        if (mk | iv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float sc = __uint_as_float(r[j]);
            if ((mk >> j) & 1u) sc += kMaskAdd;
            if ((iv >> j) & 1u) sc = -INFINITY;
            mx = fmaxf(mx, sc);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx(m_run - m_new);             // 0 on the first tile (m_run = -inf)
      // sweep 2: P = exp2(S - m) -> fp16, written over the S columns already consumed (P col = S col / 2)
      float sum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < kTile; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tb + kColS + c0, r);
        const uint32_t mk = shifted ? sm.kmask[par][my_qreg][c0 >> 5] : 0u;
        const uint32_t iv = partial ? sm.kinval[par][c0 >> 5] : 0u;
        tc::tmem_wait_ld();
        uint32_t p16[16];
        if (mk | iv) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float s0 = __uint_as_float(r[j]), s1 = __uint_as_float(r[j + 1]);
            if ((mk >> j) & 1u) s0 += kMaskAdd;
            if ((mk >> (j + 1)) & 1u) s1 += kMaskAdd;
            const float p0 = ((iv >> j) & 1u) ? 0.f : ex2_approx(s0 - m_new);
            const float p1 = ((iv >> (j + 1)) & 1u) ? 0.f : ex2_approx(s1 - m_new);
            sum += p0 + p1;
            p16[j >> 1] = pack_h2f(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_approx(__uint_as_float(r[j]) - m_new), p1 = ex2_approx(__uint_as_float(r[j + 1]) - m_new);
            sum += p0 + p1;
            p16[j >> 1] = pack_h2f(p0, p1);
          }
        }
        tc::tmem_st16(tb + kColS + c0 / 2, p16);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      // rescale the running output when some row of this warp moved its maximum (warp-collective TMEM access); PV(kt-1)
      // is complete: QK(kt) was only issued after it
      if (kt > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int c0 = 0; c0 < kC; c0 += 32) {
          uint32_t r[32];
          tc::tmem_ld32(tb + kColO + c0, r);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
          tmem_st32(tb + kColO + c0, r);
        }
      }
      tc::tmem_wait_st();
      tc::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sm.p_ready);
    }
    tc::mbar_wait(&sm.bar_o, (n_kt - 1) & 1);
    tc::tc_fence_after_sync();
    const int tok = sm.qtok[row];
    const float inv = 1.f / l_run;
#pragma unroll
    for (int c0 = 0; c0 < kC; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tb + kColO + c0, r);
      tc::tmem_wait_ld();
      if (tok >= 0) {
        float4* dst = reinterpret_cast<float4*>(out + boff + (size_t)tok * kC + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                               __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// v4: pre-packed operand tiles + TMA-unit bulk copies.  ncu of v3: the loader warps became the critical path (11 long-
// scoreboard stalls per issue): every one of the 10 query-tile CTAs of a window re-gathers and re-converts the same fp32
// K / V rows, four dependent L2 round trips per key tile.  v4 does that work ONCE per call: a pre-pass writes every
// 128-row Q / K / V tile of every window as a ready-made operand image (fp16, SWIZZLE_128B, Q pre-scaled, rows past the
// window end zero) into a caller-provided workspace; the attention kernel then feeds its shared-memory operand buffers
// with cp.async.bulk (one elected lane, mbarrier complete_tx) -- no registers, no conversions, K(t+1) in flight
// underneath softmax(t) and V(t+1) underneath QK(t+1) + softmax(t+1).
//   warps 0-3 softmax | warp 4 producer (bulk copies + shift-region bit masks) | warp 5 MMA issuer
constexpr int kPackThreads = 256;
constexpr int kV4Threads = 320;              // 8 softmax warps (2 threads per query row) + producer + MMA issuer
constexpr int kImageBytes = 2 * kBlockBytes;      // one [128][128] fp16 operand tile = two swizzled [128][64] blocks

// grid (n_tiles, B * n_windows, 3): image (b, window, tile, matrix) of the workspace
__global__ void __launch_bounds__(kPackThreads)
attn_pack_tiles_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                       unsigned char* __restrict__ ws, const WinGeomTc g) {
  __shared__ int tok_s[kTile];
  const int tid = threadIdx.x;
  const int Lw = g.wh * g.ww;
  const int n_tiles = gridDim.x;
  const int win = blockIdx.y % (g.splits * g.splits), b = blockIdx.y / (g.splits * g.splits);
  const int wy = win / g.splits, wx = win - wy * g.splits;
  const int mat = blockIdx.z;
  const float* src = (mat == 0 ? q : (mat == 1 ? k : v)) + (size_t)b * g.h * g.w * kC;
  const float scale = mat == 0 ? rsqrtf((float)kC) * 1.4426950408889634f : 1.0f;   // scores in the exp2 domain
  if (tid < kTile) {
    int tok = -1, reg = 0;
    if (blockIdx.x * kTile + tid < Lw) window_token(g, wy, wx, blockIdx.x * kTile + tid, tok, reg);
    tok_s[tid] = tok;
  }
  __syncthreads();
  unsigned char* img = ws + (((size_t)blockIdx.y * n_tiles + blockIdx.x) * 3 + mat) * kImageBytes;
  load_rows_swizzled(src, tok_s, img, scale, tid);       // same image the v2 / v3 kernels build in shared memory
}

// ---- optional timeline trace of CTA (0, 0) (debug aid, tools/attn_trace.py): lane 0 of the MMA-issuer warp, the producer
// warp and the two softmax warps of rows 0-31 record clock64() at protocol points.  Armed by mnf_debug_attn_trace(); a probe
// costs one uniform branch when not armed.
__device__ unsigned long long* g_attn_trace_buf = nullptr;
constexpr int kAttnTraceRoles = 4, kAttnTracePer = 256;
struct AttnTrace {
  unsigned long long* buf;      // this role's slice, or nullptr
  unsigned n;
  __device__ __forceinline__ void hit(int ev) {
    if (buf != nullptr && n < kAttnTracePer) buf[n++] = ((unsigned long long)clock64() << 8) | (unsigned)(ev & 255);
  }
};

struct AttnV4Smem {
  alignas(1024) unsigned char q[2][kBlockBytes];
  unsigned char k[2][kBlockBytes];
  unsigned char v[2][kBlockBytes];
  int qtok[kTile];
  uint32_t kmask[2][9][4];         // [tile parity][query region][word]: key lies in ANOTHER shift region
  uint32_t kinval[2][4];           // [tile parity][word]: key past the end of the window
  alignas(8) uint64_t bar_s;       // QK(t) complete            (tcgen05.commit)
  uint64_t bar_o;                  // PV(t) complete            (tcgen05.commit)
  uint64_t k_full;                 // K(t) (and Q, for t = 0) landed   (complete_tx)
  uint64_t v_full;                 // V(t) landed                      (complete_tx)
  uint64_t p_ready;                // P(t) in tensor memory, O rescaled   (8 softmax warps)
  uint32_t tmem_base;
  float pmax[2][2][kTile];         // [tile parity][column half][row]: row maxima of the two threads of a row
  float lsum[2][kTile];            // final exchange of the two partial row sums
};

__global__ void __launch_bounds__(kV4Threads, 2)
window_attn_tc_v4_kernel(const unsigned char* __restrict__ ws, float* __restrict__ out, const WinGeomTc g) {
  extern __shared__ unsigned char smem_dyn[];
  AttnV4Smem& sm = *reinterpret_cast<AttnV4Smem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lw = g.wh * g.ww;
  const int win = blockIdx.y % (g.splits * g.splits), b = blockIdx.y / (g.splits * g.splits);
  const int wy = win / g.splits, wx = win - wy * g.splits;
  const int q0 = blockIdx.x * kTile;
  const size_t boff = (size_t)b * g.h * g.w * kC;
  const bool shifted = (g.sh | g.sw) != 0;
  const float kLog2e = 1.4426950408889634f;
  const int n_kt = (Lw + kTile - 1) / kTile;
  const unsigned char* wbase = ws + (size_t)blockIdx.y * n_kt * 3 * kImageBytes;     // this window's images: [tile][q, k, v]

  if (tid == 0) {
    tc::mbar_init(&sm.bar_s, 1);
    tc::mbar_init(&sm.bar_o, 1);
    tc::mbar_init(&sm.k_full, 1);
    tc::mbar_init(&sm.v_full, 1);
    tc::mbar_init(&sm.p_ready, 8);
    tc::fence_mbar_init();
  }
  if (warp == 9) tc::tmem_alloc<256>(&sm.tmem_base);
  int my_qreg = 0;
  if (tid < 2 * kTile) {                      // both threads of a query row need its shift region
    int tok = -1, reg = 0;
    if (q0 + (tid & (kTile - 1)) < Lw) window_token(g, wy, wx, q0 + (tid & (kTile - 1)), tok, reg);
    if (tid < kTile) sm.qtok[tid] = tok;
    my_qreg = reg;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  AttnTrace tr{nullptr, 0u};
  {
    const int role = warp == 9 ? 0 : (warp == 8 ? 1 : (warp == 0 ? 2 : (warp == 4 ? 3 : -1)));
    unsigned long long* tb_ = g_attn_trace_buf;
    if (tb_ != nullptr && role >= 0 && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0) tr.buf = tb_ + role * kAttnTracePer;
  }

  if (warp == 8) {
    // ================================================================== producer
    for (int t = 0; t < n_kt; ++t) {
      const int par = t & 1;
      // K buffer and the mask buffers of parity `par` are free once QK(t-1) is complete (it was issued after PV(t-2),
      // which waited for softmax(t-2), the last reader of kmask[par])
      if (t > 0) tc::mbar_wait(&sm.bar_s, (t - 1) & 1);
      tr.hit(20);
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) {
        int tok = -1, reg = 0;
        if (t * kTile + w4 * 32 + lane < Lw) window_token(g, wy, wx, t * kTile + w4 * 32 + lane, tok, reg);
        if (shifted) {
#pragma unroll
          for (int r = 0; r < 9; ++r) {
            const uint32_t other = __ballot_sync(0xffffffffu, reg != r);
            if (lane == 0) sm.kmask[par][r][w4] = other;
          }
        }
        const uint32_t inval = __ballot_sync(0xffffffffu, tok < 0);
        if (lane == 0) sm.kinval[par][w4] = inval;
      }
      __syncwarp();
      if (lane == 0) {
        const unsigned char* timg = wbase + (size_t)t * 3 * kImageBytes;
        if (t == 0) {
          tc::mbar_arrive_expect_tx(&sm.k_full, 2 * kImageBytes);
          const unsigned char* qimg = wbase + (size_t)blockIdx.x * 3 * kImageBytes;
          tc::bulk_g2s(&sm.q[0][0], qimg, kBlockBytes, &sm.k_full);
          tc::bulk_g2s(&sm.q[1][0], qimg + kBlockBytes, kBlockBytes, &sm.k_full);
        } else {
          tc::mbar_arrive_expect_tx(&sm.k_full, kImageBytes);
        }
        tc::bulk_g2s(&sm.k[0][0], timg + kImageBytes, kBlockBytes, &sm.k_full);
        tc::bulk_g2s(&sm.k[1][0], timg + kImageBytes + kBlockBytes, kBlockBytes, &sm.k_full);
      }
      tr.hit(21);
      if (t > 0) tc::mbar_wait(&sm.bar_o, (t - 1) & 1);          // V buffer free: PV(t-1) complete
      tr.hit(22);
      if (lane == 0) {
        const unsigned char* timg = wbase + (size_t)t * 3 * kImageBytes;
        tc::mbar_arrive_expect_tx(&sm.v_full, kImageBytes);
        tc::bulk_g2s(&sm.v[0][0], timg + 2 * kImageBytes, kBlockBytes, &sm.v_full);
        tc::bulk_g2s(&sm.v[1][0], timg + 2 * kImageBytes + kBlockBytes, kBlockBytes, &sm.v_full);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ================================================================== MMA issuer
    const uint32_t idesc_qk = tc::umma_idesc_f16(128, 128, 0), idesc_pv = tc::umma_idesc_f16(128, 128, 1);
    for (int kt = 0; kt < n_kt; ++kt) {
      tc::mbar_wait(&sm.k_full, kt & 1);
      tr.hit(1);
      if (kt > 0) tc::mbar_wait(&sm.bar_o, (kt - 1) & 1);        // S / P columns free, O accumulation ordered
      tr.hit(2);
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = tc::umma_desc_sw128(tc::smem_u32(&sm.q[ks >> 2][0]) + (ks & 3) * 32);
          const uint64_t bd = tc::umma_desc_sw128(tc::smem_u32(&sm.k[ks >> 2][0]) + (ks & 3) * 32);
          tc::umma_ss(tmem + kColS, ad, bd, idesc_qk, ks > 0);
        }
        tc::umma_commit(&sm.bar_s);
      }
      __syncwarp();
      tr.hit(3);
      tc::mbar_wait(&sm.p_ready, kt & 1);
      tr.hit(4);
      tc::mbar_wait(&sm.v_full, kt & 1);
      tr.hit(5);
      tc::tc_fence_after_sync();
      if (tc::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // 16 keys per step: A = P columns [8 ks, 8 ks + 8), B = V rows [16 ks, 16 ks + 16)
          const uint64_t bd = tc::umma_desc_sw128_mn(tc::smem_u32(&sm.v[0][0]) + ks * 2048, kBlockBytes);
          tc::umma_ts(tmem + kColO, tmem + kColS + ks * 8, bd, idesc_pv, (kt | ks) ? 1u : 0u);
        }
        tc::umma_commit(&sm.bar_o);
      }
      __syncwarp();
      tr.hit(6);
    }
  } else {
    // ================================================================== softmax: two threads per query row
    // warp w works on TMEM lanes 32 (w & 3) .. +31 (rows) and on column half hc = w >> 2: S columns [64 hc, 64 hc + 64),
    // P columns [32 hc, 32 hc + 32), O columns [64 hc, 64 hc + 64).  The two threads of a row exchange their partial
    // row maxima through shared memory (one 256-thread named barrier per key tile, buffers alternate with the tile
    // parity); each keeps its own partial row sum until the end.
    const int hc = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float kMaskAdd = -100.0f * kLog2e;                     // transformer.py:41, :90
    float m_run = -INFINITY, l_run = 0.f;
    for (int kt = 0; kt < n_kt; ++kt) {
      const int par = kt & 1;
      const bool partial = Lw - kt * kTile < kTile;              // warp-uniform
      tc::mbar_wait(&sm.bar_s, par);
      tr.hit(10);
      tc::tc_fence_after_sync();
      float mx = -INFINITY;
#pragma unroll
      for (int c1 = 0; c1 < 64; c1 += 32) {
        const int c0 = hc * 64 + c1;
        uint32_t r[32];
        tc::tmem_ld32(tb + kColS + c0, r);
        const uint32_t mk = shifted ? sm.kmask[par][my_qreg][c0 >> 5] : 0u;
        const uint32_t iv = partial ? sm.kinval[par][c0 >> 5] : 0u;
        tc::tmem_wait_ld();
        if (mk | iv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float sc = __uint_as_float(r[j]);
            if ((mk >> j) & 1u) sc += kMaskAdd;
            if ((iv >> j) & 1u) sc = -INFINITY;
            mx = fmaxf(mx, sc);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      }
      sm.pmax[par][hc][row] = mx;
      tr.hit(11);
      asm volatile("bar.sync 2, 256;" ::: "memory");             // the 8 softmax warps
      tr.hit(12);
      mx = fmaxf(mx, sm.pmax[par][hc ^ 1][row]);
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx(m_run - m_new);             // 0 on the first tile (m_run = -inf)
      // sweep 2: P = exp2(S - m) -> fp16.  P of key columns [c, c + 32) is stored over tensor-memory columns [c / 2, c / 2 + 16):
      // the hc = 0 thread's P lands on S columns it has already consumed, but the hc = 1 thread's P (columns [32, 64)) lands on
      // S columns the hc = 0 thread still has to read in its second chunk.  So hc = 1 computes both of its chunks into
      // registers and stores them only after a barrier that hc = 0 reaches once its second chunk is in registers.
      float sum = 0.f;
      uint32_t p16[2][16];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c0 = hc * 64 + ci * 32;
        uint32_t r[32];
        tc::tmem_ld32(tb + kColS + c0, r);
        const uint32_t mk = shifted ? sm.kmask[par][my_qreg][c0 >> 5] : 0u;
        const uint32_t iv = partial ? sm.kinval[par][c0 >> 5] : 0u;
        tc::tmem_wait_ld();
        if (ci == 1 && hc == 0) asm volatile("bar.sync 3, 256;" ::: "memory");      // all of this thread's S reads are done
        if (mk | iv) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float s0 = __uint_as_float(r[j]), s1 = __uint_as_float(r[j + 1]);
            if ((mk >> j) & 1u) s0 += kMaskAdd;
            if ((mk >> (j + 1)) & 1u) s1 += kMaskAdd;
            const float p0 = ((iv >> j) & 1u) ? 0.f : ex2_approx(s0 - m_new);
            const float p1 = ((iv >> (j + 1)) & 1u) ? 0.f : ex2_approx(s1 - m_new);
            sum += p0 + p1;
            p16[ci][j >> 1] = pack_h2f(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_approx(__uint_as_float(r[j]) - m_new), p1 = ex2_approx(__uint_as_float(r[j + 1]) - m_new);
            sum += p0 + p1;
            p16[ci][j >> 1] = pack_h2f(p0, p1);
          }
        }
        if (hc == 0) tc::tmem_st16(tb + kColS + c0 / 2, p16[ci]);
      }
      if (hc == 1) {
        asm volatile("bar.sync 3, 256;" ::: "memory");
        tc::tmem_st16(tb + kColS + 32, p16[0]);
        tc::tmem_st16(tb + kColS + 48, p16[1]);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      tr.hit(13);
      if (kt > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {    // PV(kt-1) is complete: QK(kt) was only issued after it
#pragma unroll
        for (int c1 = 0; c1 < 64; c1 += 32) {
          uint32_t r[32];
          tc::tmem_ld32(tb + kColO + hc * 64 + c1, r);
          tc::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
          tmem_st32(tb + kColO + hc * 64 + c1, r);
        }
      }
      tc::tmem_wait_st();
      tc::tc_fence_before_sync();
      __syncwarp();
      tr.hit(14);
      if (lane == 0) tc::mbar_arrive(&sm.p_ready);
    }
    sm.lsum[hc][row] = l_run;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float l_tot = sm.lsum[0][row] + sm.lsum[1][row];
    tc::mbar_wait(&sm.bar_o, (n_kt - 1) & 1);
    tc::tc_fence_after_sync();
    const int tok = sm.qtok[row];
    const float inv = 1.f / l_tot;
#pragma unroll
    for (int c1 = 0; c1 < 64; c1 += 32) {
      const int c0 = hc * 64 + c1;
      uint32_t r[32];
      tc::tmem_ld32(tb + kColO + c0, r);
      tc::tmem_wait_ld();
      if (tok >= 0) {
        float4* dst = reinterpret_cast<float4*>(out + boff + (size_t)tok * kC + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                               __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused q / k / v projection + operand packing (TransformerLayer.forward, models/gmflow/transformer.py:158-163: query =
// q_proj(source), key = k_proj(target), value = v_proj(target)).  Before: three cuBLAS TF32 GEMMs wrote q / k / v as fp32
// [B, L, 128] tensors (47 MB per call at DTU size) and attn_pack_tiles_kernel re-read them to build the operand images.  Here a
// CTA = one operand image (b, window, tile, matrix): it gathers the 128 token rows of the tile from the layer's INPUT (roll /
// window partition by index arithmetic, rows past the window end zero), converts them to the fp16 swizzled A operand in shared
// memory, multiplies by the matrix' pre-swizzled fp16 weight (one 32 KB bulk copy) with 8 tcgen05.mma steps into 128 TMEM columns,
// and the epilogue writes the image (Q pre-scaled into the exp2 domain) straight into the workspace the attention kernel reads.
// fp16 operands / fp32 accumulation (the TF32 GEMMs rounded their operands to the same 10-bit mantissa).
constexpr int kProjThreads = 256;

struct ProjSmem {
  alignas(1024) unsigned char a[2][kBlockBytes];
  unsigned char w[2][kBlockBytes];
  int tok[kTile];
  alignas(8) uint64_t w_full;
  uint64_t d_full;
  uint32_t tmem_base;
};

// grid (n_tiles, B * n_windows, 3)
__global__ void __launch_bounds__(kProjThreads, 2)
attn_project_pack_kernel(const float* __restrict__ source, const float* __restrict__ target, const unsigned char* __restrict__ wpk,
                         unsigned char* __restrict__ ws, const WinGeomTc g, const int target_roll) {
  extern __shared__ unsigned char smem_dyn[];
  ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Lw = g.wh * g.ww;
  const int n_tiles = gridDim.x;
  const int win = blockIdx.y % (g.splits * g.splits), b = blockIdx.y / (g.splits * g.splits);
  const int wy = win / g.splits, wx = win - wy * g.splits;
  const int mat = blockIdx.z;
  // keys / values of batch item b come from target[(b + target_roll) % B]: FeatureTransformer's "the other view of the pair"
  // (torch.cat([x[b:], x[:b]]), transformer.py:331) without materialising the concatenation
  const int n_batch = gridDim.y / (g.splits * g.splits);
  const float* x = mat == 0 ? source + (size_t)b * g.h * g.w * kC : target + (size_t)((b + target_roll) % n_batch) * g.h * g.w * kC;
  if (tid == 0) {
    tc::mbar_init(&sm.w_full, 1);
    tc::mbar_init(&sm.d_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc<128>(&sm.tmem_base);
  if (tid < kTile) {
    int tok = -1, reg = 0;
    if (blockIdx.x * kTile + tid < Lw) window_token(g, wy, wx, blockIdx.x * kTile + tid, tok, reg);
    sm.tok[tid] = tok;
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  tc::tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(&sm.w_full, kImageBytes);
    tc::bulk_g2s(&sm.w[0][0], wpk + (size_t)mat * kImageBytes, kImageBytes, &sm.w_full);
  }
  load_rows_swizzled(x, sm.tok, &sm.a[0][0], 1.0f, tid);
  tc::fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0) {
    tc::mbar_wait(&sm.w_full, 0);
    tc::tc_fence_after_sync();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::umma_idesc_f16(128, 128);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc::umma_ss(tmem, tc::umma_desc_sw128(tc::smem_u32(&sm.a[kb][0]) + ks * 32), tc::umma_desc_sw128(tc::smem_u32(&sm.w[kb][0]) + ks * 32),
                      idesc, (kb | ks) ? 1u : 0u);
      tc::umma_commit(&sm.d_full);
    }
    __syncwarp();
  }
  tc::mbar_wait(&sm.d_full, 0);
  tc::tc_fence_after_sync();
  {
    const int q4 = warp & 3, hf = warp >> 2, row = q4 * 32 + lane;
    const float scale = mat == 0 ? rsqrtf((float)kC) * 1.4426950408889634f : 1.0f;   // scores in the exp2 domain, as attn_pack_tiles_kernel
    unsigned char* img = ws + (((size_t)blockIdx.y * n_tiles + blockIdx.x) * 3 + mat) * kImageBytes + (size_t)hf * kBlockBytes;
    const uint32_t acc = tmem + ((uint32_t)(q4 * 32) << 16) + hf * 64;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t r[16];
      tc::tmem_ld16(acc + c4 * 16, r);
      tc::tmem_wait_ld(r);
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        uint4 o;
        o.x = pack_h2f(__uint_as_float(r[h8 * 8 + 0]) * scale, __uint_as_float(r[h8 * 8 + 1]) * scale);
        o.y = pack_h2f(__uint_as_float(r[h8 * 8 + 2]) * scale, __uint_as_float(r[h8 * 8 + 3]) * scale);
        o.z = pack_h2f(__uint_as_float(r[h8 * 8 + 4]) * scale, __uint_as_float(r[h8 * 8 + 5]) * scale);
        o.w = pack_h2f(__uint_as_float(r[h8 * 8 + 6]) * scale, __uint_as_float(r[h8 * 8 + 7]) * scale);
        *reinterpret_cast<uint4*>(img + tc::sw128_offset(row, c4 * 16 + h8 * 8)) = o;
      }
    }
  }
  tc::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<128>(tmem);
}

// q_proj / k_proj / v_proj weights [128][128] fp32 (nn.Linear layout: [out][in]) -> three pre-swizzled fp16 operand images
__global__ void attn_pack_proj_weights_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                                              unsigned char* __restrict__ outp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte chunk each: 3 matrices x 2 K-blocks x 128 rows x 8 chunks
  if (i >= 3 * 2 * 128 * 8) return;
  const int mat = i >> 11, kb = (i >> 10) & 1, r = (i >> 3) & 127, c8 = i & 7;
  const float* src = (mat == 0 ? wq : (mat == 1 ? wk : wv)) + (size_t)r * kC + kb * 64 + c8 * 8;
  const float4 lo = __ldg(reinterpret_cast<const float4*>(src)), hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
  *reinterpret_cast<uint4*>(outp + (size_t)mat * kImageBytes + (size_t)kb * kBlockBytes + tc::sw128_offset(r, c8 * 8)) =
      make_uint4(pack_h2f(lo.x, lo.y), pack_h2f(lo.z, lo.w), pack_h2f(hi.x, hi.y), pack_h2f(hi.z, hi.w));
}

int64_t window_attn_tc_workspace_bytes(int B, int h, int w, int num_splits) {
  if (B <= 0 || h <= 0 || w <= 0 || num_splits <= 0 || h % num_splits || w % num_splits) return 0;
  const int64_t Lw = (int64_t)(h / num_splits) * (w / num_splits);
  const int64_t n_tiles = (Lw + kTile - 1) / kTile;
  return (int64_t)B * num_splits * num_splits * n_tiles * 3 * kImageBytes;
}

bool window_attn_tc_supports(int B, int h, int w, int C, int num_splits) {
  return C == kC && B > 0 && h % num_splits == 0 && w % num_splits == 0;
}

int launch_window_attn_tc(const float* q, const float* k, const float* v, float* out, int B, int h, int w, int C,
                          int num_splits, int with_shift, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  WinGeomTc g;
  g.h = h; g.w = w; g.splits = num_splits;
  g.wh = h / num_splits; g.ww = w / num_splits;
  g.sh = (with_shift && num_splits > 1) ? g.wh / 2 : 0;
  g.sw = (with_shift && num_splits > 1) ? g.ww / 2 : 0;
  const int Lw = g.wh * g.ww;
  static const int pipe = [] { const char* e = getenv("MNF_ATTN_PIPE"); return e ? atoi(e) : 2; }();   // A/B knob: 1 = v3 even with a workspace
  dim3 grid((Lw + kTile - 1) / kTile, B * num_splits * num_splits);
  if (pipe >= 2 && workspace && workspace_bytes >= window_attn_tc_workspace_bytes(B, h, w, num_splits) &&
      ((uintptr_t)workspace & 15) == 0) {
    const size_t smem = sizeof(AttnV4Smem) + 1024;
    static PerDevice<bool> configured_dev;
    bool& configured = configured_dev.cur();
    if (!configured) {
      MNF_CUDA_TRY(cudaFuncSetAttribute(window_attn_tc_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    dim3 pgrid(grid.x, grid.y, 3);
    attn_pack_tiles_kernel<<<pgrid, kPackThreads, 0, s>>>(q, k, v, reinterpret_cast<unsigned char*>(workspace), g);
    MNF_CUDA_TRY(cudaGetLastError());
    window_attn_tc_v4_kernel<<<grid, kV4Threads, smem, s>>>(reinterpret_cast<const unsigned char*>(workspace), out, g);
  } else {              // no workspace: every CTA gathers and converts its own operand tiles (v3)
    const size_t smem = sizeof(AttnPipeSmem) + 1024;
    static PerDevice<bool> configured_dev;
    bool& configured = configured_dev.cur();
    if (!configured) {
      MNF_CUDA_TRY(cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    window_attn_tc_kernel<<<grid, kPipeThreads, smem, s>>>(q, k, v, out, g);
  }
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

int64_t window_attn_proj_weight_bytes() { return 3 * (int64_t)kImageBytes; }

int launch_window_attn_pack_proj_weights(const float* wq, const float* wk, const float* wv, void* out, cudaStream_t s) {
  attn_pack_proj_weights_kernel<<<(3 * 2 * 128 * 8 + 255) / 256, 256, 0, s>>>(wq, wk, wv, reinterpret_cast<unsigned char*>(out));
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

// attention(q_proj(source), k_proj(target), v_proj(target)): projection + operand packing kernel, then the v4 attention kernel
int launch_window_attn_proj_tc(const float* source, const float* target, const void* proj_weights, float* out, int B, int h, int w,
                               int num_splits, int with_shift, int target_roll, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
  WinGeomTc g;
  g.h = h; g.w = w; g.splits = num_splits;
  g.wh = h / num_splits; g.ww = w / num_splits;
  g.sh = (with_shift && num_splits > 1) ? g.wh / 2 : 0;
  g.sw = (with_shift && num_splits > 1) ? g.ww / 2 : 0;
  const int Lw = g.wh * g.ww;
  if (!workspace || workspace_bytes < window_attn_tc_workspace_bytes(B, h, w, num_splits) || ((uintptr_t)workspace & 15) != 0) {
    set_error("mnf_window_attn_proj_fwd: workspace missing, misaligned or smaller than mnf_window_attn_workspace_bytes");
    return MNF_ENOMEM;
  }
  dim3 grid((Lw + kTile - 1) / kTile, B * num_splits * num_splits);
  const size_t smem = sizeof(AttnV4Smem) + 1024, psmem = sizeof(ProjSmem) + 1024;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.cur();
  if (!configured) {
    MNF_CUDA_TRY(cudaFuncSetAttribute(window_attn_tc_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA_TRY(cudaFuncSetAttribute(attn_project_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
    configured = true;
  }
  dim3 pgrid(grid.x, grid.y, 3);
  attn_project_pack_kernel<<<pgrid, kProjThreads, psmem, s>>>(source, target, reinterpret_cast<const unsigned char*>(proj_weights),
                                                              reinterpret_cast<unsigned char*>(workspace), g, ((target_roll % B) + B) % B);
  MNF_CUDA_TRY(cudaGetLastError());
  window_attn_tc_v4_kernel<<<grid, kV4Threads, smem, s>>>(reinterpret_cast<const unsigned char*>(workspace), out, g);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf

// Debug aid (tools/attn_trace.py; not part of include/matchnerf_b200.h): arm (device buffer of 4 x 256 uint64) or disarm (NULL)
// the timeline trace of the v4 attention kernel.
extern "C" int32_t mnf_debug_attn_trace(void* dev_buf) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(dev_buf);
  return cudaMemcpyToSymbol(mnf::g_attn_trace_buf, &p, sizeof(p)) == cudaSuccess ? 0 : -3;
}
