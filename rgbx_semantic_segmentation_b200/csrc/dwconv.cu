// Depthwise 3x3 (pad 1) + bias + activation on channels-last bf16 — the Mix-FFN DWConv+GELU
// (dual_segformer.py:25-33, 69-70) and the FFM ChannelEmbed DWConv+ReLU (net_utils.py:314-315),
// computed directly in the token-major layout (no NLC<->NCHW copies).  Memory-bound: each thread
// owns VEC channels x TW consecutive pixels of one image row and slides a 3 x (TW+2) register window,
// 16-byte (8-byte for VEC=4) accesses, per-CTA weights staged in shared memory (tap-major).
#include "common.cuh"
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
// Shared-memory tiled variant (the production path).  CTA = 8 rows x 32 pixels x 64 channels:
//   * the (8+2) x (32+2) halo tile of x is staged once in shared memory with 16-byte loads ([pixel][64 ch] bf16,
//     128 B per pixel => a warp reading one pixel's channel pairs is one conflict-free 128 B wavefront);
//   * thread (row r = tid/32, channel pair cp = tid%32) walks its row with a 3x3 register window (3 LDS.32 per
//     pixel), so per-channel weights, bias and - in the backward mode - the dW[9]/db accumulators are plain
//     registers (20 floats) instead of per-pixel arrays;
//   * global traffic per warp access is one full 128-byte line (64 channels x bf16), for dy reads and y/du writes.
// MODE 0: y = act(conv(x)+b)   MODE 1: du = dy * act'(conv(x)+b), dW/db reduced   MODE 2: dx = conv_flipped(du)
// ------------------------------------------------------------------------------------------------
constexpr int DT_TH = 8, DT_CH = 64;
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
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&tile[((r + i) * (TW + 2) + j) * DT_CH + cp * 2]);
          win[i][j + 1] = __bfloat1622float2(h2);
        }
      }
      bf16* orow = out + ((long)(b * H + gy) * W + x0) * ldo + c;   // advanced by ldo per pixel
#pragma unroll  // fully unrolled: the sliding-window register rotation disappears and gq[] stays in registers
      for (int px = 0; px < TW; px++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          win[i][0] = win[i][1];
          win[i][1] = win[i][2];
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&tile[((r + i) * (TW + 2) + px + 2) * DT_CH + cp * 2]);
          win[i][2] = __bfloat1622float2(h2);
        }
        const int gx = x0 + px;
        if (gx >= W) break;
        float2 a = bs;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 3; j++) ffma2(a, wt[i * 3 + j], win[i][j]);
        if (MODE == 1) {
          float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gq[px]));
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

template <int ACT, int MODE, int TW>
static void dwconv_tiled_launch_tw(const void* x, long ldx, const float* w, const float* bias, const void* dy, long lddy, void* out,
                                   long ldo, float* dw, float* db, int B, int H, int W, int C, int groups, long pgs, cudaStream_t st) {
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + DT_TH - 1) / DT_TH;
  const long ntiles = (long)B * tiles_x * tiles_y;
  const int gx = (C + DT_CH - 1) / DT_CH;
  long gy = ntiles;
  long cap = (MODE == 1 || db) ? (148L * 5 + gx - 1) / gx : (148L * 20 + gx - 1) / gx;  // reductions: few CTAs => few atomics
  cap = (cap + groups - 1) / groups;
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
static void dwconv_tiled_launch(const void* x, long ldx, const float* w, const float* bias, const void* dy, long lddy, void* out,
                                long ldo, float* dw, float* db, int B, int H, int W, int C, int groups, long pgs, cudaStream_t st) {
  if (dwconv_pick_tw(W) == 20) dwconv_tiled_launch_tw<ACT, MODE, 20>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
  else dwconv_tiled_launch_tw<ACT, MODE, 32>(x, ldx, w, bias, dy, lddy, out, ldo, dw, db, B, H, W, C, groups, pgs, st);
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
    if (flip) {
      CMX_REQUIRE(act == CMX_ACT_NONE, "dwconv3x3: flip is for the data gradient (no activation)");
      dwconv_tiled_launch<CMX_ACT_NONE, 2>(x, ldx, w, nullptr, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    } else if (act == CMX_ACT_GELU) dwconv_tiled_launch<CMX_ACT_GELU, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    else if (act == CMX_ACT_RELU) dwconv_tiled_launch<CMX_ACT_RELU, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
    else dwconv_tiled_launch<CMX_ACT_NONE, 0>(x, ldx, w, bias, nullptr, 0, y, ldy, nullptr, ysum, B, H, W, C, groups, pgs, st);
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
    if (act == CMX_ACT_GELU) dwconv_tiled_launch<CMX_ACT_GELU, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
    else if (act == CMX_ACT_RELU) dwconv_tiled_launch<CMX_ACT_RELU, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
    else dwconv_tiled_launch<CMX_ACT_NONE, 1>(x, ldx, w, bias, dy, lddy, du, lddu, dw, db, B, H, W, C, groups, pgs, st);
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
