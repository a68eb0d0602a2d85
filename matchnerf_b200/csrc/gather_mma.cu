#ifdef MNF_EXPERIMENTS   // experiment (slower than v3): compiled only with -DMNF_EXPERIMENTS, never part of the product library
// K-gather v5: the bilinear blend as a tensor-core product.
//
// Same function as gather.cu (MatchNeRF.query_cond_info, models/matchnerf.py:209-293, fused with ray casting / depth
// sampling / projection: misc/camera.py:255-286, :351-379; matchnerf.py:163-181); different machine mapping.
//
// Why: v3 (gather.cu) runs at ~49 % of the CUDA-core arithmetic roofline -- the 4-tap blend alone is 6144 FMAs per sample
// and packed fp16 math issues at half rate on this part, so no CUDA-core formulation can be more than 2x faster.  But the
// blend IS a small matrix product,  blended[item][channel] = sum_tap W[item][tap] * T[tap][channel],  with the same T for
// every ray that falls into the same neighbourhood of texels.  It runs here on the tensor cores (mma.sync.m16n8k16, fp16
// operands, fp32 accumulate):
//   * M = 16 items = 16 consecutive rays at one depth sample (adjacent pixels project 0.25 / 0.125 texel apart);
//   * K = 16 taps = a WINDOW of 8 x 2 texels (x even-aligned): it holds every bilinear cell with y0 = yw and
//     xw <= x0 <= xw + 6, i.e. normally ALL 16 rays of the batch (their cells span <= 5 columns); an item's other 12
//     weights are zero.  A ballot loop clusters the batch into windows; further windows (a batch straddling two texel
//     rows, or scattered rays) accumulate into the same product through a slow path;
//   * N = 8 channels per MMA.  The feature maps are stored x-pair interleaved (packing v4, pack.cu): one 32-bit word = the
//     same channel of texels (2i, 2i+1) = one k-pair of the B fragment, so a lane's two 16-byte loads (window rows y, y+1)
//     are the B operands of four MMAs straight from global memory, at immediate offsets from one per-view pointer;
//   * the fp32 accumulator fragments ARE the blended features (rounded once instead of four times), laid out so that a
//     lane holds 8 contiguous channels of items m and m+8: the three pair products are lane-local packed fp32 FMAs, a fine
//     cosine group (16 channels) is a lane pair and a coarse group (64 channels) a lane quad.
// v4 (first MMA version: K = 8 windows of 4 x 2 texels in a dynamic loop, A fragments parked in shared memory) was
// correct but 2.8x slower than v3: ~30 instructions of plumbing per 4 HMMAs and no L1 left (profiles/r01_ncu_summary.md).
// v5 has no window loop on the fast path: per (view, half, 32-channel block) it is 2 loads + 4 HMMAs, fully unrolled.
#include <cstdlib>

#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kRays4 = 16;        // rays per warp = M of the MMA
constexpr int kWarps4 = 2;        // warps per CTA
constexpr int kMaxWin = 16;       // windows per (batch, view, scale): at most one per ray
constexpr int kStage4 = 28;       // floats per staged row: 22 used, [22, 24) zero, 16-byte aligned rows
constexpr uint32_t kBlockB = 1024;  // bytes of one x-pair block: 256 channels x 2 texels x fp16

struct WarpSmem {
  uint4 rec[2][2 * kViews][kRays4];          // [sample of the pair][view, scale][ray]: x0, fx bits, fy bits, window id
  uint32_t win_off[2][2 * kViews][kMaxWin];  // byte offset of the window's first block (row yw, pair xw / 2)
  uint32_t win_x[2][2 * kViews][kMaxWin];    // xw
  uint32_t nwin[2][2 * kViews];
  float stage[32][kStage4];                  // conditioning rows of the 32 items of a phase
};

__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// D[16x8] (+)= A[16x16] * B[16x8], fp16 operands, fp32 accumulate.  With g = lane / 4, t = lane % 4:
//   A (row-major): a0 = (row g, k 2t..2t+1), a1 = (row g+8, k 2t..), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
//   B (col-major): b0 = (k 2t..2t+1, n g), b1 = (k 2t+8..2t+9, n g)
//   D: d0,d1 = (row g, cols 2t, 2t+1), d2,d3 = (row g+8, same cols)
__device__ __forceinline__ void mma_acc(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

// The weights item `rec` gives to the texel pair x = xw + 2t + {0, 1} of window `wi`: .x for window row yw (k = 2t..2t+1),
// .y for row yw + 1 (k = 2t+8..2t+9).  Products formed in fp32 exactly as grid_sample's bilinear weights, then rounded to
// fp16 (the same rounding v2 / v3 apply).
__device__ __forceinline__ uint2 weight_pairs(const uint4 rec, const uint32_t wi, const uint32_t xw, const int t) {
  if (rec.w != wi) return make_uint2(0u, 0u);
  const float fx = __uint_as_float(rec.y), fy = __uint_as_float(rec.z);
  const int t0 = 2 * t - ((int)rec.x - (int)xw);          // low texel relative to the cell: 0 -> 1-fx, 1 -> fx
  const float wlo = t0 == 0 ? 1.f - fx : (t0 == 1 ? fx : 0.f);
  const float whi = t0 == -1 ? 1.f - fx : (t0 == 0 ? fx : 0.f);
  const __half2 top = __floats2half2_rn(wlo * (1.f - fy), whi * (1.f - fy));
  const __half2 bot = __floats2half2_rn(wlo * fy, whi * fy);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&top), *reinterpret_cast<const uint32_t*>(&bot));
}


// 3 values x 2 items held by the lane pair (q, q^1): transpose-reduce so that even lanes keep item mrow and odd lanes item
// mrow + 8 (the coarse scale also sums over the other lane pair of the quad), then this pair's cosine (matchnerf.py:268)
template <int SC>
__device__ __forceinline__ float pair_cosine(const float2 (&acc)[2][3], const bool odd) {
  float f[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = acc[0][k].x + acc[0][k].y, hi = acc[1][k].x + acc[1][k].y;
    f[k] = (odd ? hi : lo) + __shfl_xor_sync(0xffffffffu, odd ? lo : hi, 1);
    if (SC == 0) f[k] += __shfl_xor_sync(0xffffffffu, f[k], 2);
  }
  return f[0] * rsqrtf(fmaxf(f[1], 1e-16f)) * rsqrtf(fmaxf(f[2], 1e-16f));      // max(sqrt(x), 1e-8) == sqrt(max(x, 1e-16))
}

__device__ __forceinline__ void pair_accumulate(float2 (&acc)[2][3], const float (&dA)[4][4], const float (&dB)[4][4], const bool first) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 a = make_float2(dA[j][2 * i], dA[j][2 * i + 1]);
      const float2 bb = make_float2(dB[j][2 * i], dB[j][2 * i + 1]);
      if (first && j == 0) {
        acc[i][0] = fmul2(a, bb); acc[i][1] = fmul2(a, a); acc[i][2] = fmul2(bb, bb);
      } else {
        acc[i][0] = ffma2(a, bb, acc[i][0]); acc[i][1] = ffma2(a, a, acc[i][1]); acc[i][2] = ffma2(bb, bb, acc[i][2]);
      }
    }
  }
}

// Fast path of one (batch, scale): every view needs exactly one window.  No window loop, no branches: per pair and
// 32-channel block it is 4 loads (issued one pair ahead) + 8 HMMAs + 24 packed FMAs.  A cosine group = one 32-channel
// block x a lane pair (fine scale, SC = 1) or two blocks x the lane quad (coarse scale, SC = 0).
template <int SC>
__device__ __forceinline__ void scale_fast(WarpSmem& sm, const int b, const char* fbase, const size_t mapB, const uint32_t rowB,
                                           const int mrow, const int q, const bool odd) {
  constexpr int kVA[3] = {0, 0, 1}, kHA[3] = {0, 1, 1}, kVB[3] = {1, 2, 2}, kHB[3] = {0, 0, 1};   // pairs (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1)
  constexpr int kBlk = SC ? 4 : 2, kPer = SC ? 1 : 2;
  uint32_t A[kViews][4];
  const char* P[kViews];
#pragma unroll
  for (int v = 0; v < kViews; ++v) {
    const int vs = v * 2 + SC;
    const uint32_t xw = sm.win_x[b][vs][0];
    const uint2 wa = weight_pairs(sm.rec[b][vs][mrow], 0u, xw, q), wb = weight_pairs(sm.rec[b][vs][mrow + 8], 0u, xw, q);
    A[v][0] = wa.x; A[v][1] = wb.x; A[v][2] = wa.y; A[v][3] = wb.y;
    P[v] = fbase + v * mapB + sm.win_off[b][vs][0];
  }
  // B fragments of one pair: [side A row 0, side A row 1, side B row 0, side B row 1]
  auto fetch = [&](uint4 (&L)[4], const int pr, const uint32_t goff) {
    const char* pa = P[kVA[pr]] + goff + kHA[pr] * 512;
    const char* pb = P[kVB[pr]] + goff + kHB[pr] * 512;
    L[0] = __ldg(reinterpret_cast<const uint4*>(pa));
    L[1] = __ldg(reinterpret_cast<const uint4*>(pa + rowB));
    L[2] = __ldg(reinterpret_cast<const uint4*>(pb));
    L[3] = __ldg(reinterpret_cast<const uint4*>(pb + rowB));
  };
  // two block groups per loop iteration on the fine scale: the number of (pair, block) steps per iteration is then even
  // (6 on both scales), so the double buffer of prefetched B fragments keeps its parity across iterations
  constexpr int kUn = SC ? 2 : 1, kSteps = kUn * 3 * kPer;
  uint4 L[2][4];
  fetch(L[0], 0, 0u);
#pragma unroll 1
  for (int gb0 = 0; gb0 < kBlk; gb0 += kUn) {
#pragma unroll
    for (int u = 0; u < kUn; ++u) {
      const int gb = gb0 + u;
      float simsum = 0.f;
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        float2 acc[2][3];
#pragma unroll
        for (int g = 0; g < kPer; ++g) {
          const int step = (u * 3 + pr) * kPer + g;            // position inside the loop iteration (compile time)
          const int cur = step & 1;
          {  // prefetch the next step: same iteration, or the first step of the next one (harmless wrap at the very end)
            const int nstep = (step + 1) % kSteps;
            const int nu = nstep / (3 * kPer), npr = (nstep / kPer) % 3, ng = nstep % kPer;
            const int ngb = (step + 1 == kSteps) ? (gb0 + kUn < kBlk ? gb0 + kUn : 0) : gb0 + nu;
            fetch(L[cur ^ 1], npr, (uint32_t)(ngb * kPer + ng) * 128u);
          }
          float dA[4][4], dB[4][4];
          mma_zero(dA[0], A[kVA[pr]], L[cur][0].x, L[cur][1].x);
          mma_zero(dA[1], A[kVA[pr]], L[cur][0].y, L[cur][1].y);
          mma_zero(dA[2], A[kVA[pr]], L[cur][0].z, L[cur][1].z);
          mma_zero(dA[3], A[kVA[pr]], L[cur][0].w, L[cur][1].w);
          mma_zero(dB[0], A[kVB[pr]], L[cur][2].x, L[cur][3].x);
          mma_zero(dB[1], A[kVB[pr]], L[cur][2].y, L[cur][3].y);
          mma_zero(dB[2], A[kVB[pr]], L[cur][2].z, L[cur][3].z);
          mma_zero(dB[3], A[kVB[pr]], L[cur][2].w, L[cur][3].w);
          pair_accumulate(acc, dA, dB, g == 0);
        }
        simsum += pair_cosine<SC>(acc, odd);
      }
      const float sim = simsum * (1.0f / 3.0f);              // mean over pairs, matchnerf.py:271
      const int row = b * 16 + mrow + (odd ? 8 : 0);
      if (SC == 1) sm.stage[row][2 + 2 * gb + (q >> 1)] = sim;
      else if ((q & 2) == 0) sm.stage[row][gb] = sim;
    }
  }
}

// Generic path of one (batch, scale): any number of windows per view (a batch straddling two texel rows, scattered rays
// of a training batch).  Everything is a run-time loop; A fragments are formed on the fly per window.
__device__ __noinline__ void scale_generic(WarpSmem& sm, const int b, const int sc, const char* fbase, const size_t mapB,
                                           const uint32_t rowB, const int mrow, const int q, const bool odd) {
  const int n_blk = sc ? 4 : 2, per = sc ? 1 : 2;
  auto blend = [&](float (&dd)[4][4], const int v, const int blk) {
    const int vs = v * 2 + sc;
    const int nw = (int)sm.nwin[b][vs];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) dd[j][i] = 0.f;
#pragma unroll 1
    for (int wi = 0; wi < nw; ++wi) {
      const uint32_t xw = sm.win_x[b][vs][wi];
      const uint2 wa = weight_pairs(sm.rec[b][vs][mrow], (uint32_t)wi, xw, q);
      const uint2 wb = weight_pairs(sm.rec[b][vs][mrow + 8], (uint32_t)wi, xw, q);
      const uint32_t Ax[4] = {wa.x, wb.x, wa.y, wb.y};
      const char* p = fbase + v * mapB + sm.win_off[b][vs][wi] + blk * 128;
      const uint4 B0 = __ldg(reinterpret_cast<const uint4*>(p)), B1 = __ldg(reinterpret_cast<const uint4*>(p + rowB));
      mma_acc(dd[0], Ax, B0.x, B1.x);
      mma_acc(dd[1], Ax, B0.y, B1.y);
      mma_acc(dd[2], Ax, B0.z, B1.z);
      mma_acc(dd[3], Ax, B0.w, B1.w);
    }
  };
#pragma unroll 1
  for (int gb = 0; gb < n_blk; ++gb) {
    float simsum = 0.f;
#pragma unroll 1
    for (int pr = 0; pr < 3; ++pr) {
      const int vA = pr == 2 ? 1 : 0, hA = pr == 0 ? 0 : 1;
      const int vB = pr == 0 ? 1 : 2, hB = pr == 2 ? 1 : 0;
      float2 acc[2][3];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[i][k] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int g = 0; g < per; ++g) {
        const int gq = gb * per + g;
        float dA[4][4], dB[4][4];
        blend(dA, vA, gq + 4 * hA);
        blend(dB, vB, gq + 4 * hB);
        pair_accumulate(acc, dA, dB, false);
      }
      simsum += sc ? pair_cosine<1>(acc, odd) : pair_cosine<0>(acc, odd);
    }
    const float sim = simsum * (1.0f / 3.0f);
    const int row = b * 16 + mrow + (odd ? 8 : 0);
    if (sc == 1) sm.stage[row][2 + 2 * gb + (q >> 1)] = sim;
    else if ((q & 2) == 0) sm.stage[row][gb] = sim;
  }
}

}  // namespace

__global__ void __launch_bounds__(kWarps4 * 32, 8)
gather_mma_kernel(const __grid_constant__ DevCams cams, const DevRays rays, const int S,
                  const __half* __restrict__ f0, const int h0, const int w0,
                  const __half* __restrict__ f1, const int h1, const int w1,
                  const float4* __restrict__ images, float* __restrict__ cond_f32, __half* __restrict__ cond_f16) {
  __shared__ __align__(16) WarpSmem smem_all[kWarps4];
  const uint32_t full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpSmem& sm = smem_all[wib];
  const int64_t ray0 = ((int64_t)blockIdx.x * kWarps4 + wib) * kRays4;
  if (ray0 >= rays.n_rays) return;
  const int sp = lane >> 4, rr = lane & 15;       // geometry phase: sample of the pair, ray of the warp
  const int mrow = lane >> 2, q = lane & 3;       // MMA phase: fragment row, lane of the quad
  const int64_t my_ray = min(ray0 + rr, rays.n_rays - 1);      // rays past the end repeat the last one (not stored)
  const int64_t pix = rays.points ? 0 : (rays.ray_idx ? rays.ray_idx[my_ray] : rays.first_ray + my_ray);
  float o[3], d[3];
  cast_ray(cams, pix, o, d);
  const int HW = cams.H * cams.W;
  const uint32_t wp0 = (uint32_t)(w0 + 1) >> 1, wp1 = (uint32_t)(w1 + 1) >> 1;
  const size_t map0B = (size_t)h0 * wp0 * kBlockB, map1B = (size_t)h1 * wp1 * kBlockB;
  // lane part of a B-fragment address: x-pair block q of the window row, 16-byte slot of fragment column mrow
  const uint32_t laneoff = (uint32_t)q * kBlockB + (uint32_t)mrow * 16u;
#pragma unroll
  for (int k = kCond; k < kStage4; ++k) sm.stage[lane][k] = 0.f;
  const bool odd = q & 1;

  for (int s0 = 0; s0 < S; s0 += 2) {
    // ------------------------------------------------------------ geometry phase: lane = (sample s0 + sp, ray rr)
    {
      const int s = min(s0 + sp, S - 1);
      float p[3];
      if (rays.points) {                                                         // explicit sample points (query_cond_info)
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __ldg(rays.points + ((size_t)my_ray * S + s) * 3 + i);
      } else {
        const float u = rays.jitter ? rays.jitter[my_ray * S + s] : 0.f;
        const float t = sample_depth(cams, s, S, u);
#pragma unroll
        for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(o[i], __fmul_rn(d[i], t));   // misc/camera.py:281-286
      }
      const uint32_t halfmask = 0xffffu << (16 * sp);
#pragma unroll 1
      for (int v = 0; v < kViews; ++v) {
        float uu, vv, zz;
        project_ndc(cams, v, p, uu, vv, zz);
        const float gx = __fsub_rn(__fmul_rn(uu, 2.0f), 1.0f);                   // matchnerf.py:234
        const float gy = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
        sm.stage[lane][19 + v] = (gx > -1.0f && gx < 1.0f && gy > -1.0f && gy < 1.0f) ? 1.f : 0.f;   // :248-250
        {  // colours (matchnerf.py:245)
          const float ix = grid_unnormalize(gx, cams.W), iy = grid_unnormalize(gy, cams.H);
          const float x0f = floorf(ix), y0f = floorf(iy);
          const float fx = ix - x0f, fy = iy - y0f;
          const int x0 = (int)x0f, y0 = (int)y0f;
          const int dx = x0 + 1 <= cams.W - 1 ? 1 : 0, dy = y0 + 1 <= cams.H - 1 ? cams.W : 0;
          const float4* pr = images + (size_t)v * HW + y0 * cams.W + x0;
          const float4 c00 = __ldg(pr), c01 = __ldg(pr + dx), c10 = __ldg(pr + dy), c11 = __ldg(pr + dy + dx);
          const float wa0 = (1.f - fx) * (1.f - fy), wb0 = fx * (1.f - fy), wa1 = (1.f - fx) * fy, wb1 = fx * fy;
          sm.stage[lane][10 + 3 * v + 0] = (c00.x * wa0 + c01.x * wb0) + (c10.x * wa1 + c11.x * wb1);
          sm.stage[lane][10 + 3 * v + 1] = (c00.y * wa0 + c01.y * wb0) + (c10.y * wa1 + c11.y * wb1);
          sm.stage[lane][10 + 3 * v + 2] = (c00.z * wa0 + c01.z * wb0) + (c10.z * wa1 + c11.z * wb1);
        }
#pragma unroll 1
        for (int sc = 0; sc < 2; ++sc) {
          const int vs = v * 2 + sc;
          const int w = sc ? w1 : w0, h = sc ? h1 : h0;
          const uint32_t wp = sc ? wp1 : wp0;
          const float ix = grid_unnormalize(gx, w), iy = grid_unnormalize(gy, h);
          const float x0f = floorf(ix), y0f = floorf(iy);
          const float fx = ix - x0f, fy = iy - y0f;
          const uint32_t x0u = (uint32_t)(int)x0f, y0u = (uint32_t)(int)y0f;
          // cluster the 16 cells of each half-warp's batch into 8 x 2 texel windows (x origin even); normally one
          bool covered = false;
          uint32_t win = 0, nw = 0;
          while (true) {
            const uint32_t unc = __ballot_sync(full, !covered);
            if (unc == 0) break;
            const uint32_t mine = unc & halfmask;
            const int leader = mine ? (__ffs(mine) - 1) : lane;
            const uint32_t lx = __shfl_sync(full, x0u, leader), ly = __shfl_sync(full, y0u, leader);
            // window row = the leader's; x origin = the leftmost uncovered cell of that row (whichever way the rays run)
            const uint32_t cand = (!covered && y0u == ly) ? x0u : 0xffffffffu;
            const uint32_t minx = min(__reduce_min_sync(full, sp == 0 ? cand : 0xffffffffu), 0xffffffffu);
            const uint32_t minx1 = __reduce_min_sync(full, sp == 1 ? cand : 0xffffffffu);
            if (mine) {
              const uint32_t xw = (sp ? minx1 : minx) & ~1u;
              if (!covered && y0u == ly && x0u - xw <= 6u) { covered = true; win = nw; }
              if (lane == leader) {
                sm.win_off[sp][vs][nw] = (ly * wp + (xw >> 1)) * kBlockB;
                sm.win_x[sp][vs][nw] = xw;
              }
              ++nw;
            }
          }
          if (rr == 0) sm.nwin[sp][vs] = nw;
          sm.rec[sp][vs][rr] = make_uint4(x0u, __float_as_uint(fx), __float_as_uint(fy), win);
        }
      }
    }
    __syncwarp();
    // ------------------------------------------------------------ gather phase: the (up to) two 16-ray batches
    const int n_b = min(2, S - s0);
    for (int b = 0; b < n_b; ++b) {
      const char* fb0 = reinterpret_cast<const char*>(f0) + laneoff;
      const char* fb1 = reinterpret_cast<const char*>(f1) + laneoff;
      if (sm.nwin[b][0] + sm.nwin[b][2] + sm.nwin[b][4] == 3u) scale_fast<0>(sm, b, fb0, map0B, wp0 * kBlockB, mrow, q, odd);
      else scale_generic(sm, b, 0, fb0, map0B, wp0 * kBlockB, mrow, q, odd);
      if (sm.nwin[b][1] + sm.nwin[b][3] + sm.nwin[b][5] == 3u) scale_fast<1>(sm, b, fb1, map1B, wp1 * kBlockB, mrow, q, odd);
      else scale_generic(sm, b, 1, fb1, map1B, wp1 * kBlockB, mrow, q, odd);
    }
    __syncwarp();
    // ------------------------------------------------------------ store: the 2 samples of a ray are contiguous
    if (cond_f16) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int idx = p * 32 + lane;
        const int ray = idx >> 3, bsel = (idx >> 2) & 1, part = idx & 3;
        if (ray0 + ray < rays.n_rays && bsel < n_b) {
          uint4 pk = make_uint4(0u, 0u, 0u, 0u);
          if (part < 3) {
            const float4 x = *reinterpret_cast<const float4*>(&sm.stage[bsel * 16 + ray][part * 8]);
            const float4 y = *reinterpret_cast<const float4*>(&sm.stage[bsel * 16 + ray][part * 8 + 4]);
            const __half2 a0 = __floats2half2_rn(x.x, x.y), a1 = __floats2half2_rn(x.z, x.w);
            const __half2 a2 = __floats2half2_rn(y.x, y.y), a3 = __floats2half2_rn(y.z, y.w);
            pk.x = *reinterpret_cast<const uint32_t*>(&a0); pk.y = *reinterpret_cast<const uint32_t*>(&a1);
            pk.z = *reinterpret_cast<const uint32_t*>(&a2); pk.w = *reinterpret_cast<const uint32_t*>(&a3);
          }
          const size_t n = (size_t)(ray0 + ray) * S + s0 + bsel;
          *reinterpret_cast<uint4*>(cond_f16 + n * kCondPad + part * 8) = pk;
        }
      }
    }
    if (cond_f32) {
#pragma unroll
      for (int p = 0; p < 11; ++p) {
        const int idx = p * 32 + lane;              // 16 rays x 22 float2 (2 samples x 22 floats = 176 contiguous bytes per ray)
        const int ray = idx / 22, e = 2 * (idx - ray * 22);
        const int bsel = e >= kCond ? 1 : 0, k = e - bsel * kCond;
        if (ray0 + ray < rays.n_rays && bsel < n_b)
          *reinterpret_cast<float2*>(cond_f32 + ((size_t)(ray0 + ray) * S + s0) * kCond + e) =
              *reinterpret_cast<const float2*>(&sm.stage[bsel * 16 + ray][k]);
      }
    }
    __syncwarp();
  }
}

int launch_gather_mma(const DevCams& cams, const DevRays& rays, int S, const __half* f0, int h0, int w0,
                      const __half* f1, int h1, int w1, const float* images, float* cond_f32, __half* cond_f16,
                      cudaStream_t s) {
  if (rays.n_rays <= 0) return MNF_OK;
  if ((int64_t)h0 * ((w0 + 1) / 2) >= (1 << 22) || (int64_t)h1 * ((w1 + 1) / 2) >= (1 << 22)) {
    set_error("feature map too large for 32-bit block offsets");
    return MNF_EUNSUPPORTED;
  }
  const int64_t groups = (rays.n_rays + kRays4 - 1) / kRays4;
  const int64_t blocks = (groups + kWarps4 - 1) / kWarps4;
  gather_mma_kernel<<<(unsigned)blocks, kWarps4 * 32, 0, s>>>(cams, rays, S, f0, h0, w0, f1, h1, w1,
                                                            reinterpret_cast<const float4*>(images), cond_f32, cond_f16);
  MNF_CUDA_TRY(cudaGetLastError());
  return MNF_OK;
}

}  // namespace mnf

#endif  // MNF_EXPERIMENTS
