// Depthwise 3x3 (pad 1) + bias + activation on channels-last bf16 - the Mix-FFN DWConv+GELU
// (dual_segformer.py:25-33, 69-70) and the FFM ChannelEmbed DWConv+ReLU (net_utils.py:314-315),
// computed directly in the token-major layout (no NLC<->NCHW copies).  Three generations live here:
//   1. dwconv_fwd_kernel / dwconv_bwd_pre_kernel: register-window kernels reading global memory directly
//      (CMX_DWCONV_LEGACY=1, single-group launches only; kept as the simplest correct form for A/B runs);
//   2. dwconv_tiled_kernel: halo tile staged by cp.async, thread = (row, channel pair), packed FFMA2
//      (CMX_DWCONV_TMA=0, or tensors a tensor map cannot address);
//   3. dwconv_tma_kernel: the PRODUCTION path - the same arithmetic, tiles delivered by TMA with a zero-filled halo,
//      two stages on mbarriers, dy through shared memory (see the comment above it).
// All are issue / FP32-pipe bound, not HBM bound (profiles/r2_ncu_full_dwconv_cpasync_pipes.txt).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>
#include <stdlib.h>
extern std::atomic<long long> g_cmx_launches;

template <int ACT>
__device__ __forceinline__ float act_f(float u) {
  if (ACT == CMX_ACT_RELU) return fmaxf(u, 0.f);
  if (ACT == CMX_ACT_GELU) return gelu_f(u);
  return u;
}
template <int ACT>
__device__ __forceinline__ float act_grad_f(float u) {
  if (ACT == CMX_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (ACT == CMX_ACT_GELU) return gelu_grad_f(u);
  return 1.f;
}

template <int VEC> struct VecIO;
template <> struct VecIO<8> {
  static __device__ __forceinline__ void ld(const bf16* p, float* f) { load8(p, f); }
  static __device__ __forceinline__ void st(bf16* p, const float* f) { store8(p, f); }
};
template <> struct VecIO<4> {
  static __device__ __forceinline__ void ld(const bf16* p, float* f) { load4(p, f); }
  static __device__ __forceinline__ void st(bf16* p, const float* f) { store4(p, f); }
};

constexpr int DW_CG = 32;  // channel groups per CTA (threadIdx.x)
constexpr int DW_PY = 8;   // pixel strips per CTA (threadIdx.y)

// ------------------------------------------------------------------------------------------------
// forward / data-gradient (FLIP) kernel
// ------------------------------------------------------------------------------------------------
template <int ACT, bool FLIP, int TW>
__global__ void __launch_bounds__(DW_CG* DW_PY) dwconv_fwd_kernel(const bf16* __restrict__ x, long ldx, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, bf16* __restrict__ y, long ldy,
                                                                  int B, int H, int W, int C) {
  pdl_trigger();
  constexpr int VEC = 8;
  __shared__ __align__(16) float sw[9][DW_CG * VEC];
  __shared__ __align__(16) float sb[DW_CG * VEC];
  const int cbase = blockIdx.x * DW_CG * VEC;
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < 9 * DW_CG * VEC; i += DW_CG * DW_PY) {
    const int tap = i / (DW_CG * VEC), cl = i % (DW_CG * VEC);
    const int c = cbase + cl;
    const int t = FLIP ? 8 - tap : tap;
    sw[tap][cl] = c < C ? w[c * 9 + t] : 0.f;
  }
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < DW_CG * VEC; i += DW_CG * DW_PY) {
    const int c = cbase + i;
    sb[i] = (bias && c < C) ? bias[c] : 0.f;
  }
  __syncthreads();
  const int cl = threadIdx.x * VEC;
  const int c = cbase + cl;
  if (c >= C) return;
  const int strips_w = (W + TW - 1) / TW;
  const long nstrips = (long)B * H * strips_w;
  for (long s = (long)blockIdx.y * DW_PY + threadIdx.y; s < nstrips; s += (long)gridDim.y * DW_PY) {
    const int xs = (int)(s % strips_w) * TW;
    const long t = s / strips_w;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    float acc[TW][VEC];
#pragma unroll
    for (int p = 0; p < TW; p++)
#pragma unroll
      for (int i = 0; i < VEC; i++) acc[p][i] = sb[cl + i];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const int y2 = yy + dy;
      if (y2 < 0 || y2 >= H) continue;
      const bf16* rowp = x + ((long)(b * H + y2) * W) * ldx + c;
      float in[TW + 2][VEC];
#pragma unroll
      for (int j = 0; j < TW + 2; j++) {
        const int x2 = xs + j - 1;
        if (x2 >= 0 && x2 < W) VecIO<VEC>::ld(rowp + (long)x2 * ldx, in[j]);
        else {
#pragma unroll
          for (int i = 0; i < VEC; i++) in[j][i] = 0.f;
        }
      }
#pragma unroll
      for (int dx = 0; dx < 3; dx++) {
        float wv[VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) wv[i] = sw[(dy + 1) * 3 + dx][cl + i];
#pragma unroll
        for (int p = 0; p < TW; p++)
#pragma unroll
          for (int i = 0; i < VEC; i++) acc[p][i] = fmaf(wv[i], in[p + dx][i], acc[p][i]);
      }
    }
#pragma unroll
    for (int p = 0; p < TW; p++) {
      const int x2 = xs + p;
      if (x2 < W) {
        float o[VEC];
#pragma unroll
        for (int i = 0; i < VEC; i++) o[i] = act_f<ACT>(acc[p][i]);
        VecIO<VEC>::st(y + ((long)(b * H + yy) * W + x2) * ldy + c, o);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Shared-memory tiled variant (the production path of round 1; now the fallback).  CTA = 8 rows x TW pixels x 64 channels:
//   * the (8+2) x (32+2) halo tile of x is staged once in shared memory with 16-byte loads ([pixel][64 ch] bf16,
//     128 B per pixel => a warp reading one pixel's channel pairs is one conflict-free 128 B wavefront);
//   * thread (row r = tid/32, channel pair cp = tid%32) walks its row with a 3x3 register window (3 LDS.32 per
//     pixel), so per-channel weights, bias and - in the backward mode - the dW[9]/db accumulators are plain
//     registers (20 floats) instead of per-pixel arrays;
//   * global traffic per warp access is one full 128-byte line (64 channels x bf16), for dy reads and y/du writes.
// MODE 0: y = act(conv(x)+b)   MODE 1: du = dy * act'(conv(x)+b), dW/db reduced   MODE 2: dx = conv_flipped(du)
// ------------------------------------------------------------------------------------------------
constexpr int DT_TH = 8, DT_CH = 64;
// bf16 pair -> float2 in two integer ops (shift, mask); __bfloat1622float2 compiles to three (PRMT + 2 shifts) and the kernels
// below unpack 3-4 pairs per pixel
__device__ __forceinline__ float2 bf2_unpack(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
// tile width TW: 32 pixels, or 20 for the narrow late-stage maps (W = 40 / 20 / 80 ...: a 32-wide tile would be 62 % full at W = 40)
constexpr int dt_min_ctas(int mode, int tw) { return mode == 1 ? 2 : (tw <= 20 ? 4 : 3); }

template <int ACT, int MODE, int TW>
__global__ void __launch_bounds__(256, dt_min_ctas(MODE, TW)) dwconv_tiled_kernel(const bf16* __restrict__ x, long ldx, const float* __restrict__ w,
                                                           const float* __restrict__ bias, const bf16* __restrict__ dy, long lddy,
                                                           bf16* __restrict__ out, long ldo, float* __restrict__ dw,
                                                           float* __restrict__ db, int B, int H, int W, int C, int tiles_x,
                                                           int tiles_y, long ntiles, long pgs) {
  pdl_trigger();
  if (gridDim.z > 1) {  // grouped launch: group g = samples [g*B, (g+1)*B) of the stacked tensors; every parameter of
    const long g = blockIdx.z;  // group g lies g * pgs ELEMENTS behind the given pointer (flat parameter / gradient buffer)
    const long rows = g * B * H * W;
    x += rows * ldx; out += rows * ldo; w += g * pgs;
    if (dy) dy += rows * lddy;
    if (bias) bias += g * pgs;
    if (dw) dw += g * pgs;
    if (db) db += g * pgs;
  }
  __shared__ __align__(16) bf16 tile[(DT_TH + 2) * (TW + 2) * DT_CH];  // 43.5 KB at TW = 32, 27.5 KB at TW = 20
  __shared__ float sred[20][DT_CH];
  const int tid = threadIdx.x;
  const int r = tid >> 5, cp = tid & 31;
  const int cbase = blockIdx.x * DT_CH;
  const int c = cbase + cp * 2;
  const bool c_ok = c < C;  // C is even
  float2 wt[9], bs = make_float2(0.f, 0.f);  // .x = channel c, .y = channel c + 1: every FMA below is one packed FFMA2
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const int tt = MODE == 2 ? 8 - t : t;
    wt[t].x = c_ok ? w[(long)c * 9 + tt] : 0.f;
    wt[t].y = c_ok ? w[(long)(c + 1) * 9 + tt] : 0.f;
  }
  if (MODE != 2 && bias && c_ok) { bs.x = bias[c]; bs.y = bias[c + 1]; }
  float2 gw[9], gb = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; t++) gw[t] = make_float2(0.f, 0.f);

  for (long tl = blockIdx.y; tl < ntiles; tl += gridDim.y) {
    const int tx = (int)(tl % tiles_x);
    const int ty = (int)((tl / tiles_x) % tiles_y);
    const int b = (int)(tl / ((long)tiles_x * tiles_y));
    const int x0 = tx * TW, y0 = ty * DT_TH;
    __syncthreads();  // previous tile fully consumed
    // ---- stage the halo tile: (TH+2)*(TW+2) pixels x 8 chunks of 8 channels, as asynchronous 16-byte copies
    // (cp.async with zero fill outside the image): all ~11 copies of a thread are in flight together
    {
      constexpr int NCH = (DT_TH + 2) * (TW + 2) * (DT_CH / 8);
      const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
#pragma unroll
      for (int k = 0; k < (NCH + 255) / 256; k++) {
        const int i = tid + k * 256;
        if (i < NCH) {
          const int ch8 = i & 7;
          const int pix = i >> 3;
          const int px = pix % (TW + 2), py = pix / (TW + 2);
          const int gy = y0 + py - 1, gx = x0 + px - 1;
          const bool inb = gy >= 0 && gy < H && gx >= 0 && gx < W && cbase + ch8 * 8 < C;
          const bf16* src = inb ? x + ((long)(b * H + gy) * W + gx) * ldx + cbase + ch8 * 8 : x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + (uint32_t)(pix * DT_CH + ch8 * 8) * 2u), "l"(src),
                       "r"(inb ? 16 : 0)
                       : "memory");
        }
      }
    }
    const int gy = y0 + r;
    // MODE 1: this thread's 32 dy values of the row are requested before waiting for the tile (independent loads)
    uint32_t gq[TW];
    if (MODE == 1) {
      // (one 64-bit address per row, then a pointer increment per pixel: the per-pixel index arithmetic was ~10 % of the
      // kernel's instructions and the kernel is issue bound)
      const bf16* dyp = dy + ((long)(b * H + (gy < H ? gy : 0)) * W + x0) * lddy + c;
#pragma unroll
      for (int px = 0; px < TW; px++) {
        const int gx = x0 + px;
        gq[px] = 0u;
        if (gy < H && c_ok && gx < W) gq[px] = __ldg(reinterpret_cast<const unsigned int*>(dyp));
        dyp += lddy;
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (gy < H && c_ok) {
      // 3x3 window: win[row][col] for the 2 channels; columns slide
      float2 win[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
          win[i][j + 1] = bf2_unpack(*reinterpret_cast<const uint32_t*>(&tile[((r + i) * (TW + 2) + j) * DT_CH + cp * 2]));
        }
      }
      bf16* orow = out + ((long)(b * H + gy) * W + x0) * ldo + c;   // advanced by ldo per pixel
#pragma unroll  // fully unrolled: the sliding-window register rotation disappears and gq[] stays in registers
      for (int px = 0; px < TW; px++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          win[i][0] = win[i][1];
          win[i][1] = win[i][2];
          win[i][2] = bf2_unpack(*reinterpret_cast<const uint32_t*>(&tile[((r + i) * (TW + 2) + px + 2) * DT_CH + cp * 2]));
        }
        const int gx = x0 + px;
        if (gx >= W) break;
        float2 a = bs;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) ffma2(a, wt[i * 3 + j], win[i][j]);
        if (MODE == 1) {
          float2 g = bf2_unpack(gq[px]);
          if (ACT == CMX_ACT_GELU) {
            g = fmul2(g, gelu_grad2(a));
          } else {
            g.x *= act_grad_f<ACT>(a.x);
            g.y *= act_grad_f<ACT>(a.y);
          }
          gb.x += g.x;
          gb.y += g.y;
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) ffma2(gw[i * 3 + j], g, win[i][j]);
          *reinterpret_cast<__nv_bfloat162*>(orow) = __floats2bfloat162_rn(g.x, g.y);
        } else {
          if (ACT == CMX_ACT_GELU) a = gelu2(a);
          else { a.x = act_f<ACT>(a.x); a.y = act_f<ACT>(a.y); }
          const __nv_bfloat162 o2 = __floats2bfloat162_rn(a.x, a.y);
          *reinterpret_cast<__nv_bfloat162*>(orow) = o2;
          if (db) {  // per-channel sum of what was written (bias gradient of the layer that produced x's gradient)
            const float2 of = __bfloat1622float2(o2);
            gb.x += of.x;
            gb.y += of.y;
          }
        }
        orow += ldo;
      }
    }
  }
  if (MODE != 1 && db) {
    __syncthreads();
    for (int i = tid; i < DT_CH; i += 256) sred[9][i] = 0.f;
    __syncthreads();
    if (c_ok) {
      atomicAdd(&sred[9][cp * 2], gb.x);
      atomicAdd(&sred[9][cp * 2 + 1], gb.y);
    }
    __syncthreads();
    for (int i = tid; i < DT_CH; i += 256)
      if (cbase + i < C) atomicAdd(db + cbase + i, sred[9][i]);
  }
  if (MODE == 1) {
    // reduce the 8 row-threads of every channel pair, then one atomicAdd per (channel, tap) per CTA
    __syncthreads();
    for (int i = tid; i < 20 * DT_CH; i += 256) (&sred[0][0])[i] = 0.f;
    __syncthreads();
    if (c_ok) {
#pragma unroll
      for (int t = 0; t < 9; t++) {
        atomicAdd(&sred[t][cp * 2], gw[t].x);
        atomicAdd(&sred[t][cp * 2 + 1], gw[t].y);
      }
      atomicAdd(&sred[9][cp * 2], gb.x);
      atomicAdd(&sred[9][cp * 2 + 1], gb.y);
    }
    __syncthreads();
    for (int i = tid; i < 10 * DT_CH; i += 256) {
      const int t = i / DT_CH, l = i % DT_CH;
      const int cc = cbase + l;
      if (cc < C) {
        if (t < 9) atomicAdd(dw + (long)cc * 9 + t, sred[t][l]);
        else if (db) atomicAdd(db + cc, sred[9][l]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant (the production path).  Same thread mapping and arithmetic as dwconv_tiled_kernel, but
//   * the (8+2) x (TW+2) halo tile of x - and in MODE 1 the 8 x TW tile of dy - arrive by ONE cp.async.bulk.tensor each from a
//     rank-4 tensor map [C, W, H, images]: the image border (coordinates -1 / W / H) and channels >= C are zero-filled by the
//     copy engine, so the ~25 instructions per 16-byte chunk of index arithmetic and bounds tests of the cp.async loop are gone
//     (ncu, stage 1: they were 28 % of the executed instructions of a kernel that is issue / FP32-pipe bound);
//   * two stages: the copy of tile i+1 is in flight while tile i is computed (mbarrier complete_tx), so the load latency is
//     hidden inside one CTA instead of across co-resident CTAs;
//   * dy is read from shared memory (one LDS per pixel) instead of 32 prefetched registers: 96 instead of 128 registers.
// ------------------------------------------------------------------------------------------------
template <int MODE, int TW>
struct DtmCfg {
  static constexpr uint32_t XB = (uint32_t)(DT_TH + 2) * (TW + 2) * DT_CH * 2;
  static constexpr uint32_t DB = MODE == 1 ? (uint32_t)DT_TH * TW * DT_CH * 2 : 0u;
  static constexpr uint32_t STAGE = XB + DB;             // multiples of 128 B
  static constexpr uint32_t SMEM = 2 * STAGE + 128;      // + alignment slack
  static constexpr int MIN_CTAS = MODE == 1 ? 2 : 4;
};

template <int ACT, int MODE, int TW>
__global__ void __launch_bounds__(256, (DtmCfg<MODE, TW>::MIN_CTAS)) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmX,
                                                           const __grid_constant__ CUtensorMap tmDY, const float* __restrict__ w,
                                                           const float* __restrict__ bias, bf16* __restrict__ out, long ldo,
                                                           float* __restrict__ dw, float* __restrict__ db, int B, int H, int W, int C,
                                                           int tiles_x, int tiles_y, long ntiles, long pgs) {
  using Cfg = DtmCfg<MODE, TW>;
  pdl_trigger();
  extern __shared__ uint8_t dsm_raw[];
  __shared__ float sred[20][DT_CH];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t sbase = (smem_u32(dsm_raw) + 127u) & ~127u;
  const uint32_t bar0 = smem_u32(&bars[0]);
  const long g = blockIdx.z;   // group: samples [g*B, (g+1)*B) of the stacked tensors, parameters g * pgs elements further
  out += g * B * H * W * ldo; w += g * pgs;
  if (bias) bias += g * pgs;
  if (dw) dw += g * pgs;
  if (db) db += g * pgs;
  const int tid = threadIdx.x;
  const int r = tid >> 5, cp = tid & 31;
  const int cbase = blockIdx.x * DT_CH;
  const int c = cbase + cp * 2;
  const bool c_ok = c < C;  // C is even
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    if (MODE == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDY)) : "memory");
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float2 wt[9], bs = make_float2(0.f, 0.f);  // .x = channel c, .y = channel c + 1: every FMA below is one packed FFMA2
#pragma unroll
  for (int t = 0; t < 9; t++) {
    const int tt = MODE == 2 ? 8 - t : t;
    wt[t].x = c_ok ? w[(long)c * 9 + tt] : 0.f;
    wt[t].y = c_ok ? w[(long)(c + 1) * 9 + tt] : 0.f;
  }
  if (MODE != 2 && bias && c_ok) { bs.x = bias[c]; bs.y = bias[c + 1]; }
  float2 gw[9], gb = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 9; t++) gw[t] = make_float2(0.f, 0.f);
  __syncthreads();

  const int per_img = tiles_x * tiles_y;
  auto issue = [&](long tl, int stage) {   // thread 0 only
    const int tx = (int)(tl % tiles_x);
    const int ty = (int)((tl / tiles_x) % tiles_y);
    const int img = (int)(g * B + tl / per_img);
    const uint32_t bar = bar0 + 8u * (uint32_t)stage;
    const uint32_t dst = sbase + (uint32_t)stage * Cfg::STAGE;
    mbar_expect_tx(bar, Cfg::STAGE);
    tma_load_4d(dst, &tmX, bar, cbase, tx * TW - 1, ty * DT_TH - 1, img);
    if (MODE == 1) tma_load_4d(dst + Cfg::XB, &tmDY, bar, cbase, tx * TW, ty * DT_TH, img);
  };
  if (tid == 0 && (long)blockIdx.y < ntiles) issue(blockIdx.y, 0);
  int it = 0;
  for (long tl = blockIdx.y; tl < ntiles; tl += gridDim.y, it++) {
    const int stage = it & 1;
    // the other stage was last read in iteration it-1, which ended with a __syncthreads
    if (tid == 0 && tl + gridDim.y < ntiles) issue(tl + gridDim.y, stage ^ 1);
    const int tx = (int)(tl % tiles_x);
    const int ty = (int)((tl / tiles_x) % tiles_y);
    const int b = (int)(tl / per_img);
    const int x0 = tx * TW, y0 = ty * DT_TH;
    const int gy = y0 + r;
    mbar_wait(bar0 + 8u * (uint32_t)stage, (uint32_t)(it >> 1) & 1u);
    if (gy < H && c_ok) {
      // shared-memory addresses: x tile [(TH+2)][(TW+2)][64 ch], dy tile [TH][TW][64 ch], 128 B per pixel
      const uint32_t xs = sbase + (uint32_t)stage * Cfg::STAGE + (uint32_t)(r * (TW + 2) * DT_CH + cp * 2) * 2u;
      const uint32_t ds = sbase + (uint32_t)stage * Cfg::STAGE + Cfg::XB + (uint32_t)(r * TW * DT_CH + cp * 2) * 2u;
      auto lds32 = [](uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; };
      float2 win[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) win[i][j + 1] = bf2_unpack(lds32(xs + (uint32_t)((i * (TW + 2) + j) * DT_CH) * 2u));
      bf16* orow = out + ((long)(b * H + gy) * W + x0) * ldo + c;   // advanced by ldo per pixel
#pragma unroll
      for (int px = 0; px < TW; px++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          win[i][0] = win[i][1];
          win[i][1] = win[i][2];
          win[i][2] = bf2_unpack(lds32(xs + (uint32_t)((i * (TW + 2) + px + 2) * DT_CH) * 2u));
        }
        if (x0 + px >= W) break;
        // three independent row chains (3 deep) + two adds instead of one 9-deep FFMA2 chain: the kernel stalls on
        // fixed-latency dependencies with 16-32 warps per SM
        float2 a = bs, a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 3; j++) {
          ffma2(a, wt[j], win[0][j]);
          ffma2(a1, wt[3 + j], win[1][j]);
          ffma2(a2, wt[6 + j], win[2][j]);
        }
        a = fadd2(a, fadd2(a1, a2));
        if (MODE == 1) {
          float2 gq = bf2_unpack(lds32(ds + (uint32_t)(px * DT_CH) * 2u));
          if (ACT == CMX_ACT_GELU) {
            gq = fmul2(gq, gelu_grad2(a));
          } else {
            gq.x *= act_grad_f<ACT>(a.x);
            gq.y *= act_grad_f<ACT>(a.y);
          }
          gb = fadd2(gb, gq);
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) ffma2(gw[i * 3 + j], gq, win[i][j]);
          *reinterpret_cast<__nv_bfloat162*>(orow) = __floats2bfloat162_rn(gq.x, gq.y);
        } else {
          if (ACT == CMX_ACT_GELU) a = gelu2(a);
          else { a.x = act_f<ACT>(a.x); a.y = act_f<ACT>(a.y); }
          const __nv_bfloat162 o2 = __floats2bfloat162_rn(a.x, a.y);
          *reinterpret_cast<__nv_bfloat162*>(orow) = o2;
          if (db) {  // per-channel sum of what was written (bias gradient of the layer that produced x's gradient)
            const float2 of = __bfloat1622float2(o2);
            gb.x += of.x;
            gb.y += of.y;
          }
        }
        orow += ldo;
      }
    }
    __syncthreads();  // every thread is done with this stage: it may be refilled at the top of the next iteration
  }
  if (MODE != 1 && db) {
    for (int i = tid; i < DT_CH; i += 256) sred[9][i] = 0.f;
    __syncthreads();
    if (c_ok) {
      atomicAdd(&sred[9][cp * 2], gb.x);
      atomicAdd(&sred[9][cp * 2 + 1], gb.y);
    }
    __syncthreads();
    for (int i = tid; i < DT_CH; i += 256)
      if (cbase + i < C) atomicAdd(db + cbase + i, sred[9][i]);
  }
  if (MODE == 1) {
    // reduce the 8 row-threads of every channel pair, then one atomicAdd per (channel, tap) per CTA
    for (int i = tid; i < 20 * DT_CH; i += 256) (&sred[0][0])[i] = 0.f;
    __syncthreads();
    if (c_ok) {
#pragma unroll
      for (int t = 0; t < 9; t++) {
        atomicAdd(&sred[t][cp * 2], gw[t].x);
        atomicAdd(&sred[t][cp * 2 + 1], gw[t].y);
      }
      atomicAdd(&sred[9][cp * 2], gb.x);
      atomicAdd(&sred[9][cp * 2 + 1], gb.y);
    }
    __syncthreads();
    for (int i = tid; i < 10 * DT_CH; i += 256) {
      const int t = i / DT_CH, l = i % DT_CH;
      const int cc = cbase + l;
      if (cc < C) {
        if (t < 9) atomicAdd(dw + (long)cc * 9 + t, sred[t][l]);
        else if (db) atomicAdd(db + cc, sred[9][l]);
      }
    }
  }
}

// rank-4 bf16 tensor map [C, W, H, images] over a token-major [images*H*W, ld] tensor, box [64, bw, bh, 1], no swizzle, zero fill
static int dwconv_make_map(CUtensorMap* tm, const void* ptr, long ld, int C, int W, int H, long images, int bw, int bh) {
  PFN_cmxEncodeTiled enc = cmx_get_encode();
  if (!enc) { cmx_set_error("cuTensorMapEncodeTiled unavailable"); return -2; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)images};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
  cuuint32_t box[4] = {(cuuint32_t)DT_CH, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { cmx_set_error("dwconv3x3: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
  return 0;
}

template <int ACT, int MODE, int TW>
static int dwconv_tma_launch_tw(const void* x, long ldx, const float* w, const float* bias, const void* dy, long lddy, void* out,
                                long ldo, float* dw, float* db, int B, int H, int W, int C, int groups, long pgs, cudaStream_t st) {
  using Cfg = DtmCfg<MODE, TW>;
  CUtensorMap tmX, tmDY;
  int rc = dwconv_make_map(&tmX, x, ldx, C, W, H, (long)groups * B, TW + 2, DT_TH + 2);
  if (rc) return rc;
  if (MODE == 1) {
    rc = dwconv_make_map(&tmDY, dy, lddy, C, W, H, (long)groups * B, TW, DT_TH);
    if (rc) return rc;
  } else {
    tmDY = tmX;
  }
  auto kern = dwconv_tma_kernel<ACT, MODE, TW>;
  static thread_local PerDeviceOnce once;
  static thread_local int occ = 1, sms = 148;
  if (once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute(dwconv_tma): %s", cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, Cfg::SMEM);
    if (occ < 1) occ = 1;
    once.mark();
  }
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + DT_TH - 1) / DT_TH;
  const long ntiles = (long)B * tiles_x * tiles_y;
  const int gx = (C + DT_CH - 1) / DT_CH;
  static int waves_env = -1;
  if (waves_env < 0) {
    const char* e = getenv("CMX_DWCONV_WAVES");
    waves_env = e ? atoi(e) : 0;
  }
  const int waves = waves_env > 0 ? waves_env : ((MODE == 1 || db) ? 1 : 2);   // reducing kernels: one wave = fewest atomics
  long gy = (long)sms * occ * waves / ((long)gx * groups);
  if (gy < 1) gy = 1;
  if (gy > ntiles) gy = ntiles;
  dim3 grid(gx, (unsigned)gy, (unsigned)groups);
  kern<<<grid, 256, Cfg::SMEM, st>>>(tmX, tmDY, w, bias, (bf16*)out, ldo, dw, db, B, H, W, C, tiles_x, tiles_y, ntiles, pgs);
  return 0;
}

// tile width of the TMA path: 16 or 20 pixels, whichever stages fewer pixel columns (padded width + 2 halo columns per tile)
static int dwconv_tma_pick_tw(int W) {
  const long c16 = (long)((W + 15) / 16) * 18, c20 = (long)((W + 19) / 20) * 22;
  return c20 < c16 ? 20 : 16;
}
static bool dwconv_tma_ok(const void* x, long ldx, const void* dy, long lddy, int C) {
  static int off = -1;
  if (off < 0) off = getenv("CMX_DWCONV_TMA") != nullptr && getenv("CMX_DWCONV_TMA")[0] == '0' ? 1 : 0;
  if (off) return false;
  if (C % 8 || ldx % 8 || (((uintptr_t)x) & 15)) return false;          // tensor-map strides / base: multiples of 16 bytes
  if (dy && (lddy % 8 || (((uintptr_t)dy) & 15))) return false;
  return true;
}

template <int ACT, int MODE, int TW>
static void dwconv_tiled_launch_tw(const void* x, long ldx, const float* w, const float* bias, const void* dy, long lddy, void* out,
                                   long ldo, float* dw, float* db, int B, int H, int W, int C, int groups, long pgs, cudaStream_t st) {
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + DT_TH - 1) / DT_TH;
  const long ntiles = (long)B * tiles_x * tiles_y;
  const int gx = (C + DT_CH - 1) / DT_CH;
  long gy = ntiles;
  // grid = whole waves of resident CTAs (every CTA walks ntiles / gy tiles of equal cost): the earlier fixed caps (5 or 20 CTAs
  // per SM) left a 2.5-wave grid for the 2-resident-CTA backward kernel, i.e. a last wave with half the SMs idle.
  // CMX_DWCONV_WAVES: waves per launch (default 1 for the reducing kernels - fewest atomics -, 2 otherwise); 0 = the old caps.
  static int waves_env = -1;
  if (waves_env < 0) {
    const char* e = getenv("CMX_DWCONV_WAVES");
    waves_env = e ? atoi(e) : 100;
  }
  long cap;
  if (waves_env == 0) {
    cap = (MODE == 1 || db) ? (148L * 5 + gx - 1) / gx : (148L * 20 + gx - 1) / gx;  // reductions: few CTAs => few atomics
    cap = (cap + groups - 1) / groups;
  } else {
    static thread_local PerDeviceOnce occ_once;
    static thread_local int occ = 0, sms = 148;
    if (occ_once.pending()) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwconv_tiled_kernel<ACT, MODE, TW>, 256, 0);
      if (occ < 1) occ = 1;
      occ_once.mark();
    }
    const int waves = waves_env == 100 ? ((MODE == 1 || db) ? 1 : 2) : waves_env;
    cap = (long)sms * occ * waves / ((long)gx * groups);
    if (cap < 1) cap = 1;
  }
  if (gy > cap) gy = cap;
  dim3 grid(gx, (unsigned)gy, (unsigned)groups);
  dwconv_tiled_kernel<ACT, MODE, TW><<<grid, 256, 0, st>>>((const bf16*)x, ldx, w, bias, (const bf16*)dy, lddy, (bf16*)out, ldo, dw, db,
                                                           B, H, W, C, tiles_x, tiles_y, ntiles, pgs);
}

// tile width with the fewest staged pixel columns (padded width + 2 halo columns per tile); ties go to the wider tile
static int dwconv_pick_tw(int W) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("CMX_DWCONV_TW");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 32 || forced == 20) return forced;
  const long c32 = (long)((W + 31) / 32) * 34, c20 = (long)((W + 19) / 20) * 22;
  return c20 < c32 ? 20 : 32;
}

template <int ACT, int MODE>
static int dwconv_tiled_launch(const void* x, long ldx, const float* w, const float* bias, const void* dy, long lddy, void* out,
                               long ldo, float* dw, float* db, int B, int H, int W, int C, int groups, long pgs, cudaStream_t st) {
  if (dwconv_tma_ok(x, ldx, MODE == 1 ? dy : nullptr, lddy, C)) {
    if (dwconv_tma_pick_tw(W) == 20) return dwconv_tma_launch_tw<ACT, MODE, 20>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
    return dwconv_tma_launch_tw<ACT, MODE, 16>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
  }
  if (dwconv_pick_tw(W) == 20) dwconv_tiled_launch_tw<ACT, MODE, 20>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
  else dwconv_tiled_launch_tw<ACT, MODE, 32>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
  return 0;
}

CMX_API int cmx_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w, const float* bias, int act, int flip, void* y,
                              int64_t ldy, float* ysum, int B, int H, int W, int C, int groups, int64_t param_gs, void* stream) {
  CMX_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "dwconv3x3: C and ld must be multiples of 8 (C=%d)", C);
  CMX_REQUIRE(groups >= 1 && groups <= 65535, "dwconv3x3: groups=%d", groups);
  cudaStream_t st = (cudaStream_t)stream;
  if ((long)B * H * W == 0) return 0;
  const long pgs = param_gs;
  CMX_REQUIRE(groups == 1 || getenv("CMX_DWCONV_LEGACY") == nullptr, "dwconv3x3: grouped launches need the tiled kernel");
  if (getenv("CMX_DWCONV_LEGACY") == nullptr) {
    int rc;
    if (flip) {
      CMX_REQUIRE(act == CMX_ACT_NONE, "dwconv3x3: flip is for the data gradient (no activation)");
      rc = dwconv_tiled_launch<CMX_ACT_NONE, 2>(x, ldx, w, nullptr, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    } else if (act == CMX_ACT_GELU) rc = dwconv_tiled_launch<CMX_ACT_GELU, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    else if (act == CMX_ACT_RELU) rc = dwconv_tiled_launch<CMX_ACT_RELU, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    else rc = dwconv_tiled_launch<CMX_ACT_NONE, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    if (rc) return rc;
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("dwconv3x3_tiled");
    return 0;
  }
  CMX_REQUIRE(!ysum, "dwconv3x3: ysum is only available on the tiled path");
  constexpr int TW = 4;
  const long nstrips = (long)B * H * ((W + TW - 1) / TW);
  if (nstrips == 0) return 0;
  const int gx = cdiv(C, DW_CG * 8);
  long gy = (nstrips + DW_PY - 1) / DW_PY;
  const long cap = (148L * 16 + gx - 1) / gx;  // ~16 CTAs per SM worth of work in flight, grid-stride beyond
  if (gy > cap) gy = cap;
  dim3 grid(gx, (unsigned)gy), block(DW_CG, DW_PY);
#define DW_L(ACTv, FLIPv) \
  dwconv_fwd_kernel<ACTv, FLIPv, TW><<<grid, block, 0, st>>>((const bf16*)x, ldx, w, bias, (bf16*)y, ldy, B, H, W, C)
  if (flip) { CMX_REQUIRE(act == CMX_ACT_NONE, "dwconv3x3: flip is for the data gradient (no activation)"); DW_L(CMX_ACT_NONE, true); }
  else if (act == CMX_ACT_GELU) DW_L(CMX_ACT_GELU, false);
  else if (act == CMX_ACT_RELU) DW_L(CMX_ACT_RELU, false);
  else DW_L(CMX_ACT_NONE, false);
#undef DW_L
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("dwconv3x3_fwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// backward, part 1: recompute u = conv(x)+b, du = dy * act'(u) (written bf16), and reduce
// dW[c,tap] = sum du * x(p+tap), db[c] = sum du.  VEC = 4 channels, TW = 2 pixels per thread.
// ------------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(DW_CG* DW_PY) dwconv_bwd_pre_kernel(const bf16* __restrict__ x, long ldx, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, const bf16* __restrict__ dy,
                                                                      long lddy, bf16* __restrict__ du, long lddu,
                                                                      float* __restrict__ dw, float* __restrict__ db, int B, int H,
                                                                      int W, int C) {
  pdl_trigger();
  constexpr int VEC = 4, TW = 2;
  __shared__ __align__(16) float sw[9][DW_CG * VEC];
  __shared__ __align__(16) float sb[DW_CG * VEC];
  __shared__ float sacc[10][DW_CG * VEC];
  const int cbase = blockIdx.x * DW_CG * VEC;
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < 9 * DW_CG * VEC; i += DW_CG * DW_PY) {
    const int tap = i / (DW_CG * VEC), cl = i % (DW_CG * VEC);
    const int c = cbase + cl;
    sw[tap][cl] = c < C ? w[c * 9 + tap] : 0.f;
  }
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < 10 * DW_CG * VEC; i += DW_CG * DW_PY) (&sacc[0][0])[i] = 0.f;
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < DW_CG * VEC; i += DW_CG * DW_PY) {
    const int c = cbase + i;
    sb[i] = (bias && c < C) ? bias[c] : 0.f;
  }
  __syncthreads();
  const int cl = threadIdx.x * VEC;
  const int c = cbase + cl;
  float gw[9][VEC], gb[VEC];
#pragma unroll
  for (int t = 0; t < 9; t++)
#pragma unroll
    for (int i = 0; i < VEC; i++) gw[t][i] = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; i++) gb[i] = 0.f;

  if (c < C) {
    const int strips_w = (W + TW - 1) / TW;
    const long nstrips = (long)B * H * strips_w;
    for (long s = (long)blockIdx.y * DW_PY + threadIdx.y; s < nstrips; s += (long)gridDim.y * DW_PY) {
      const int xs = (int)(s % strips_w) * TW;
      const long t = s / strips_w;
      const int yy = (int)(t % H);
      const int b = (int)(t / H);
      float in[3][TW + 2][VEC];
      float acc[TW][VEC];
#pragma unroll
      for (int p = 0; p < TW; p++)
#pragma unroll
        for (int i = 0; i < VEC; i++) acc[p][i] = sb[cl + i];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const int y2 = yy + r - 1;
        const bool rok = (y2 >= 0 && y2 < H);
        const bf16* rowp = x + ((long)(b * H + (rok ? y2 : 0)) * W) * ldx + c;
#pragma unroll
        for (int j = 0; j < TW + 2; j++) {
          const int x2 = xs + j - 1;
          if (rok && x2 >= 0 && x2 < W) VecIO<VEC>::ld(rowp + (long)x2 * ldx, in[r][j]);
          else {
#pragma unroll
            for (int i = 0; i < VEC; i++) in[r][j][i] = 0.f;
          }
        }
#pragma unroll
        for (int dx = 0; dx < 3; dx++)
#pragma unroll
          for (int p = 0; p < TW; p++)
#pragma unroll
            for (int i = 0; i < VEC; i++) acc[p][i] = fmaf(sw[r * 3 + dx][cl + i], in[r][p + dx][i], acc[p][i]);
      }
#pragma unroll
      for (int p = 0; p < TW; p++) {
        const int x2 = xs + p;
        if (x2 < W) {
          const long row = (long)(b * H + yy) * W + x2;
          float g[VEC];
          VecIO<VEC>::ld(dy + row * lddy + c, g);
#pragma unroll
          for (int i = 0; i < VEC; i++) {
            g[i] *= act_grad_f<ACT>(acc[p][i]);
            gb[i] += g[i];
          }
          VecIO<VEC>::st(du + row * lddu + c, g);
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++)
#pragma unroll
              for (int i = 0; i < VEC; i++) gw[r * 3 + dx][i] = fmaf(g[i], in[r][p + dx][i], gw[r * 3 + dx][i]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
      for (int i = 0; i < VEC; i++) atomicAdd(&sacc[t][cl + i], gw[t][i]);
#pragma unroll
    for (int i = 0; i < VEC; i++) atomicAdd(&sacc[9][cl + i], gb[i]);
  }
  __syncthreads();
  for (int i = threadIdx.y * DW_CG + threadIdx.x; i < 10 * DW_CG * VEC; i += DW_CG * DW_PY) {
    const int t = i / (DW_CG * VEC), l = i % (DW_CG * VEC);
    const int cc = cbase + l;
    if (cc < C) {
      if (t < 9) atomicAdd(dw + cc * 9 + t, sacc[t][l]);
      else if (db) atomicAdd(db + cc, sacc[9][l]);
    }
  }
}

CMX_API int cmx_dwconv3x3_bwd_pre(const void* x, int64_t ldx, const float* w, const float* bias, int act, const void* dy,
                                  int64_t lddy, void* du, int64_t lddu, float* dw, float* db, int B, int H, int W, int C,
                                  int groups, int64_t param_gs, void* stream) {
  CMX_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && lddu % 4 == 0, "dwconv3x3_bwd_pre: C/ld %% 4");
  CMX_REQUIRE(groups >= 1 && groups <= 65535, "dwconv3x3_bwd_pre: groups=%d", groups);
  cudaStream_t st = (cudaStream_t)stream;
  if ((long)B * H * W == 0) return 0;
  const long pgs = param_gs;
  const bool tiled = getenv("CMX_DWCONV_LEGACY") == nullptr && C % 8 == 0 && ldx % 8 == 0;
  CMX_REQUIRE(groups == 1 || tiled, "dwconv3x3_bwd_pre: grouped launches need the tiled kernel (C %% 8 == 0)");
  if (tiled) {
    int rc;
    if (act == CMX_ACT_GELU) rc = dwconv_tiled_launch<CMX_ACT_GELU, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
    else if (act == CMX_ACT_RELU) rc = dwconv_tiled_launch<CMX_ACT_RELU, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
    else rc = dwconv_tiled_launch<CMX_ACT_NONE, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
    if (rc) return rc;
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("dwconv3x3_bwd_pre_tiled");
    return 0;
  }
  const long nstrips = (long)B * H * ((W + 1) / 2);
  if (nstrips == 0) return 0;
  const int gx = cdiv(C, DW_CG * 4);
  long gy = (nstrips + DW_PY - 1) / DW_PY;
  const long cap = (148L * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  dim3 grid(gx, (unsigned)gy), block(DW_CG, DW_PY);
#define DW_B(ACTv) \
  dwconv_bwd_pre_kernel<ACTv><<<grid, block, 0, st>>>((const bf16*)x, ldx, w, bias, (const bf16*)dy, lddy, (bf16*)du, lddu, dw, db, B, H, W, C)
  if (act == CMX_ACT_GELU) DW_B(CMX_ACT_GELU);
  else if (act == CMX_ACT_RELU) DW_B(CMX_ACT_RELU);
  else DW_B(CMX_ACT_NONE);
#undef DW_B
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("dwconv3x3_bwd_pre");
  return 0;
}
