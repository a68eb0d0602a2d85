// K-gather v4: the bilinear blend as a tensor-core product.
//
// Same function as gather.cu (MatchNeRF.query_cond_info, models/matchnerf.py:209-293, fused with ray casting / depth
// sampling / projection: misc/camera.py:255-286, :351-379; matchnerf.py:163-181); different machine mapping.
//
// Why: the ncu capture of v3 (profiles/r01_ncu_summary.md) shows the kernel bound by the fp16 FMA pipe -- HFMA2 and
// FHFMA issue at half rate on this part (ncu's own peak: hfma x4 == ffma x2 thread-ops per clock), and the blend
// `sum_tap w_tap * texel_tap` costs 4 HFMA2 per channel pair per (sample, view, scale): 384 of the ~1460 warp
// instructions per 4 samples, 768 of the ~2160 cycles.  But the blend IS a small matrix product,
//     blended[item][channel] = sum_tap W[item][tap] * T[tap][channel],
// with the same T for every ray that falls into the same neighbourhood of texels.  v4 runs it on the tensor cores
// (mma.sync.m16n8k8, fp16 operands, fp32 accumulate):
//   * M = 16 items = 16 consecutive rays at one depth sample (adjacent pixels project 0.25 / 0.125 texel apart);
//   * K = 8 taps = a WINDOW of 4 x 2 texels (x even-aligned); an item's 2 x 2 bilinear cell lies inside a window, its
//     other 4 weights are zero.  The items of a batch are clustered into windows with a ballot loop (usually 1-3
//     windows per view and scale); the product accumulates over windows;
//   * N = 8 channels per MMA.  The feature maps are stored x-pair interleaved (packing v4, pack.cu): one 32-bit word =
//     the same channel of texels (2i, 2i+1) = one k-pair of the B fragment, so a lane's 16-byte load is the B operand
//     of four MMAs, straight from global memory (8 lanes x 16 B = one full 128 B line per window block);
//   * the fp32 accumulator fragments ARE the blended features (rounded once instead of four times), laid out so that a
//     lane holds 8 contiguous channels of items m and m+8: the three pair products are lane-local packed fp32 FMAs, a
//     fine cosine group (16 channels) is a lane pair and a coarse group (64 channels) a lane quad.
// Per item this is ~55 MMA-pipe cycles and ~200 issue slots instead of ~540 cycles of fp16-pipe-bound CUDA-core work.
#include <cstdlib>

#include "mnf_common.cuh"

namespace mnf {

namespace {

constexpr int kRays4 = 16;        // rays per warp = M of the MMA
constexpr int kWarps4 = 2;        // warps per CTA
constexpr int kMaxWin = 16;       // windows per (batch, view, scale): at most one per ray
constexpr int kFragWin = 8;       // windows whose A fragments are cached in shared memory (more: recomputed on the fly)
constexpr int kStage4 = 28;       // floats per staged row: 22 used, [22, 24) zero, 16-byte aligned rows
constexpr uint32_t kBlockB = 1024;  // bytes of one x-pair block: 256 channels x 2 texels x fp16

struct WarpSmem {
  uint4 rec[2][2 * kViews][kRays4];          // [sample of the pair][view, scale][ray]: x0, fx bits, fy bits, window id
  uint32_t win_off[2][2 * kViews][kMaxWin];  // byte offset of the window's first block (row yw, pair xw / 2)
  uint32_t win_x[2][2 * kViews][kMaxWin];    // xw
  uint32_t nwin[2][2 * kViews];
  uint2 afrag[kViews][kFragWin][32];         // lane-private A fragments of the current (batch, scale)
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

// D[16x8] += A[16x8] * B[8x8]: A row-major fp16 (a0: row lane/4, a1: row lane/4 + 8; k = 2*(lane%4), +1),
// B column-major fp16 (k = 2*(lane%4), +1; n = lane/4), D fp32 (d0,d1: row lane/4, cols 2*(lane%4), +1; d2,d3: row + 8)
__device__ __forceinline__ void mma_16x8x8(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

// One half of an A-fragment register pair: the weights item `rec` gives to the two texels (k = 2q, 2q+1) this lane
// covers in window `wi` (x origin xw).  q = 2*row + pair: texels x = xw + 2*(q&1) + {0, 1} of window row q >> 1.
__device__ __forceinline__ uint32_t weight_pair(const uint4 rec, const uint32_t wi, const uint32_t xw, const int q) {
  if (rec.w != wi) return 0u;
  const float fx = __uint_as_float(rec.y), fy = __uint_as_float(rec.z);
  const int t0 = 2 * (q & 1) - ((int)rec.x - (int)xw);     // low texel relative to the cell: 0 -> 1-fx, 1 -> fx
  const float wy = (q >> 1) ? fy : 1.f - fy;
  const float wlo = t0 == 0 ? 1.f - fx : (t0 == 1 ? fx : 0.f);
  const float whi = t0 == -1 ? 1.f - fx : (t0 == 0 ? fx : 0.f);
  const __half2 h = __floats2half2_rn(wlo * wy, whi * wy);
  return *reinterpret_cast<const uint32_t*>(&h);
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
  // lane part of a B-fragment address: block (window row q >> 1, pair q & 1), 16-byte slot of fragment column mrow
  const uint32_t laneoff0 = ((uint32_t)(q >> 1) * wp0 + (uint32_t)(q & 1)) * kBlockB + (uint32_t)mrow * 16u;
  const uint32_t laneoff1 = ((uint32_t)(q >> 1) * wp1 + (uint32_t)(q & 1)) * kBlockB + (uint32_t)mrow * 16u;
#pragma unroll
  for (int k = kCond; k < kStage4; ++k) sm.stage[lane][k] = 0.f;

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
          // cluster the 16 cells of each half-warp's batch into 4 x 2 texel windows (x origin even)
          bool covered = false;
          uint32_t win = 0, nw = 0;
          while (true) {
            const uint32_t unc = __ballot_sync(full, !covered);
            if (unc == 0) break;
            const uint32_t mine = unc & halfmask;
            const int leader = mine ? (__ffs(mine) - 1) : lane;
            const uint32_t lx = __shfl_sync(full, x0u, leader), ly = __shfl_sync(full, y0u, leader);
            if (mine) {
              const uint32_t xw = lx & ~1u;
              if (!covered && y0u == ly && x0u - xw <= 2u) { covered = true; win = nw; }
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
#pragma unroll 1
      for (int sc = 0; sc < 2; ++sc) {
        const char* fbase = reinterpret_cast<const char*>(sc ? f1 : f0) + (sc ? laneoff1 : laneoff0);
        const size_t mapB = sc ? map1B : map0B;
        // A fragments (rows mrow, mrow + 8) of every window of the three views, kept in lane-private shared memory
#pragma unroll 1
        for (int v = 0; v < kViews; ++v) {
          const int vs = v * 2 + sc;
          const int nw = min((int)sm.nwin[b][vs], kFragWin);
          const uint4 r0 = sm.rec[b][vs][mrow], r1 = sm.rec[b][vs][mrow + 8];
          for (int wi = 0; wi < nw; ++wi) {
            const uint32_t xw = sm.win_x[b][vs][wi];
            sm.afrag[v][wi][lane] = make_uint2(weight_pair(r0, wi, xw, q), weight_pair(r1, wi, xw, q));
          }
        }
        // blended features of one (view, half) for 32-channel block gq: dd[j] = channels 32*gq + 8*q + {j, 4 + j}
        auto blend_side = [&](float (&dd)[4][4], const int v, const int half, const int gq) {
          const int vs = v * 2 + sc;
          const int nw = (int)sm.nwin[b][vs];
          const char* base = fbase + v * mapB + (gq + 4 * half) * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) dd[j][i] = 0.f;
          uint4 B = __ldg(reinterpret_cast<const uint4*>(base + sm.win_off[b][vs][0]));
#pragma unroll 1
          for (int wi = 0; wi < nw; ++wi) {
            uint2 af;
            if (wi < kFragWin) {
              af = sm.afrag[v][wi][lane];
            } else {
              const uint32_t xw = sm.win_x[b][vs][wi];
              af = make_uint2(weight_pair(sm.rec[b][vs][mrow], wi, xw, q), weight_pair(sm.rec[b][vs][mrow + 8], wi, xw, q));
            }
            const uint4 Bc = B;
            if (wi + 1 < nw) B = __ldg(reinterpret_cast<const uint4*>(base + sm.win_off[b][vs][wi + 1]));
            mma_16x8x8(dd[0], af.x, af.y, Bc.x);
            mma_16x8x8(dd[1], af.x, af.y, Bc.y);
            mma_16x8x8(dd[2], af.x, af.y, Bc.z);
            mma_16x8x8(dd[3], af.x, af.y, Bc.w);
          }
        };
        // a cosine group = one 32-channel block x a lane pair (fine scale, 16 channels per half-block) or two blocks
        // x the lane quad (coarse scale, 64 channels)
        const int n_blk = sc ? 4 : 2, gq_per_blk = sc ? 1 : 2;
        const bool odd = q & 1;
#pragma unroll 1
        for (int gb = 0; gb < n_blk; ++gb) {
          float simsum = 0.f;
#pragma unroll 1
          for (int pr = 0; pr < 3; ++pr) {
            // pairs (v0h0,v1h0) (v0h1,v2h0) (v1h1,v2h1): every (view, half) is blended exactly once per channel block
            const int vA = pr == 2 ? 1 : 0, hA = pr == 0 ? 0 : 1;
            const int vB = pr == 0 ? 1 : 2, hB = pr == 2 ? 1 : 0;
            float2 acc[2][3];        // [item mrow / mrow + 8][<A,B>, <A,A>, <B,B>], packed over the column pair
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int k = 0; k < 3; ++k) acc[i][k] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int g = 0; g < gq_per_blk; ++g) {
              const int gq = gb * gq_per_blk + g;
              float dA[4][4], dB[4][4];
              blend_side(dA, vA, hA, gq);
              blend_side(dB, vB, hB, gq);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const float2 a = make_float2(dA[j][2 * i], dA[j][2 * i + 1]);
                  const float2 bb = make_float2(dB[j][2 * i], dB[j][2 * i + 1]);
                  acc[i][0] = ffma2(a, bb, acc[i][0]);
                  acc[i][1] = ffma2(a, a, acc[i][1]);
                  acc[i][2] = ffma2(bb, bb, acc[i][2]);
                }
              }
            }
            // lanes q and q^1 hold the two halves of a 16-channel run: transpose-reduce so that even lanes keep item
            // mrow and odd lanes item mrow + 8; the coarse scale also sums over the other lane pair of the quad
            float f[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float lo = acc[0][k].x + acc[0][k].y, hi = acc[1][k].x + acc[1][k].y;
              f[k] = (odd ? hi : lo) + __shfl_xor_sync(full, odd ? lo : hi, 1);
              if (sc == 0) f[k] += __shfl_xor_sync(full, f[k], 2);
            }
            const float na = fmaxf(sqrtf(f[1]), 1e-8f), nb = fmaxf(sqrtf(f[2]), 1e-8f);      // matchnerf.py:268
            simsum += __fdividef(f[0], na * nb);
          }
          const float sim = simsum * (1.0f / 3.0f);                                        // mean over pairs, :271
          const int row = b * 16 + mrow + (odd ? 8 : 0);
          if (sc == 1) sm.stage[row][2 + 2 * gb + (q >> 1)] = sim;
          else if ((q & 2) == 0) sm.stage[row][gb] = sim;
        }
      }
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
