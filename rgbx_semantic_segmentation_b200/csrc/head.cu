// Decoder-side streaming kernels: fused bilinear-upsample + sum of the four per-stage embeddings
// (MLPDecoder.py:66-77 with linear_fuse applied per stage at native resolution — the 2048-channel
// concat is never materialised), its adjoint, the single-pass upsample x4 + cross-entropy loss with
// in-kernel gradient scatter (builder.py:233,249), eval logits upsample, and the confusion matrix /
// argmax metric (utils/metric.py:8-15).
#include "common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>
extern std::atomic<long long> g_cmx_launches;

#define LAUNCH_DONE(name)      \
  do {                         \
    g_cmx_launches++;          \
    CMX_CHECK_LAUNCH(name);    \
    return 0;                  \
  } while (0)

// ---- out = bias + z0 + sum_i bilinear(z_i) ---------------------------------------------------------------
struct UpSrc {
  const bf16* z;
  int H, W;
  float sh, sw;  // in/out scale (fp32, as PyTorch computes it)
};
// thread = NG 8-channel groups (stride C/8/NG, so every load instruction of a warp stays contiguous) of one output pixel:
// the index decode, the three bilinear source positions and the 12 corner addresses are paid once per NG groups (the
// one-group form is issue bound: 666 instructions per thread, 78 % issue-active at 2 TB/s, ncu), interpolation in packed
// fp32x2 FMAs
template <typename TO, typename I, int NG>
__global__ void __launch_bounds__(256) upsample_sum_kernel(const bf16* __restrict__ z0, UpSrc s1, UpSrc s2, UpSrc s3,
                                                           const float* __restrict__ bias, TO* __restrict__ out, int B, int H0,
                                                           int W0, int C) {
  pdl_trigger();
  const int tpp = (C >> 3) / NG;  // threads per pixel
  const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;  // I: unsigned when the element count allows (cheap divisions)
  const I total = (I)B * H0 * W0 * tpp;
  if (idx >= total) return;
  const int j = (int)(idx % (I)tpp);
  const I pix = idx / (I)tpp;
  const int x = (int)(pix % (I)W0);
  const int y = (int)((pix / (I)W0) % (I)H0);
  const int b = (int)(pix / ((I)W0 * H0));
  const UpSrc srcs[3] = {s1, s2, s3};
  const bf16* corner[3][4];
  float2 wgt[3][4];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const UpSrc& s = srcs[k];
    if (!s.z) continue;
    int y0, y1, x0, x1;
    float ly, lx;
    bilin_src(y, s.sh, s.H, y0, y1, ly);
    bilin_src(x, s.sw, s.W, x0, x1, lx);
    const bf16* base = s.z + (long)b * s.H * s.W * C;
    corner[k][0] = base + ((long)y0 * s.W + x0) * C;
    corner[k][1] = base + ((long)y0 * s.W + x1) * C;
    corner[k][2] = base + ((long)y1 * s.W + x0) * C;
    corner[k][3] = base + ((long)y1 * s.W + x1) * C;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    wgt[k][0] = make_float2(w00, w00);
    wgt[k][1] = make_float2(w01, w01);
    wgt[k][2] = make_float2(w10, w10);
    wgt[k][3] = make_float2(w11, w11);
  }
  const bf16* p0 = z0 + (long)pix * C;
  TO* po = out + (long)pix * C;
#pragma unroll
  for (int gi = 0; gi < NG; gi++) {
    const int c = (j + gi * tpp) * 8;
    float2 acc[4];
    if (bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + c), b1 = *reinterpret_cast<const float4*>(bias + c + 4);
      acc[0] = make_float2(b0.x, b0.y); acc[1] = make_float2(b0.z, b0.w);
      acc[2] = make_float2(b1.x, b1.y); acc[3] = make_float2(b1.z, b1.w);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] = make_float2(0.f, 0.f);
    }
    {
      const uint4 u = *reinterpret_cast<const uint4*>(p0 + c);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 f = __bfloat1622float2(h[i]);
        acc[i].x += f.x;
        acc[i].y += f.y;
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (!srcs[k].z) continue;
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; q++) u[q] = *reinterpret_cast<const uint4*>(corner[k][q] + c);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
#pragma unroll
        for (int i = 0; i < 4; i++) ffma2(acc[i], wgt[k][q], __bfloat1622float2(h[i]));
      }
    }
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; i++) { o[2 * i] = acc[i].x; o[2 * i + 1] = acc[i].y; }
    store8(po + c, o);
  }
}
CMX_API int cmx_upsample_sum_fwd(const void* z0, const void* z1, const void* z2, const void* z3, int H0, int W0, int H1, int W1,
                                 int H2, int W2, int H3, int W3, const float* bias, void* out, int out_dtype, int B, int C,
                                 void* stream) {
  CMX_REQUIRE(C % 8 == 0, "upsample_sum: C %% 8");
  const long total = (long)B * H0 * W0 * (C >> 3);
  if (total == 0) return 0;
  UpSrc s1{(const bf16*)z1, H1, W1, (float)H1 / (float)H0, (float)W1 / (float)W0};
  UpSrc s2{(const bf16*)z2, H2, W2, (float)H2 / (float)H0, (float)W2 / (float)W0};
  UpSrc s3{(const bf16*)z3, H3, W3, (float)H3 / (float)H0, (float)W3 / (float)W0};
  const int ng = (C >> 3) % 4 == 0 ? 4 : 1;
  const long nthreads = total / ng;
#define UPS(TO, I)                                                                                                          \
  do {                                                                                                                      \
    if (ng == 4)                                                                                                            \
      upsample_sum_kernel<TO, I, 4><<<cdiv(nthreads, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)z0, s1, s2, s3, bias, \
                                                                                           (TO*)out, B, H0, W0, C);          \
    else                                                                                                                    \
      upsample_sum_kernel<TO, I, 1><<<cdiv(nthreads, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)z0, s1, s2, s3, bias, \
                                                                                           (TO*)out, B, H0, W0, C);          \
  } while (0)
  const bool small = total < (1L << 31);
  if (out_dtype == CMX_F32) {
    if (small) UPS(float, unsigned);
    else UPS(float, long);
  } else {
    if (small) UPS(bf16, unsigned);
    else UPS(bf16, long);
  }
#undef UPS
  LAUNCH_DONE("upsample_sum_fwd");
}

// ---- adjoint of one bilinear source (gather over the output pixels whose stencil touches (yi,xi)) -------
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const bf16* __restrict__ dout, int Ho, int Wo, bf16* __restrict__ dz, int Hi,
                                                           int Wi, int B, int C, float sh, float sw) {
  pdl_trigger();
  const int c8 = C >> 3;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * Hi * Wi * c8;
  if (idx >= total) return;
  const int c = (int)(idx % c8) * 8;
  const long pix = idx / c8;
  const int xi = (int)(pix % Wi);
  const int yi = (int)((pix / Wi) % Hi);
  const int b = (int)(pix / ((long)Wi * Hi));
  // candidate output range: src = (dst+0.5)*s-0.5 in (yi-1, yi+1)  =>  dst in ((yi-0.5)/s-0.5, (yi+1.5)/s-0.5), +-1 slack
  const float rh = 1.f / sh, rw = 1.f / sw;
  int ylo = (int)floorf((yi - 0.5f) * rh - 0.5f) - 1, yhi = (int)ceilf((yi + 1.5f) * rh - 0.5f) + 1;
  int xlo = (int)floorf((xi - 0.5f) * rw - 0.5f) - 1, xhi = (int)ceilf((xi + 1.5f) * rw - 0.5f) + 1;
  if (ylo < 0) ylo = 0;
  if (xlo < 0) xlo = 0;
  if (yhi > Ho - 1) yhi = Ho - 1;
  if (xhi > Wo - 1) xhi = Wo - 1;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  for (int y = ylo; y <= yhi; y++) {
    int y0, y1;
    float ly;
    bilin_src(y, sh, Hi, y0, y1, ly);
    const float wy = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
    if (wy == 0.f) continue;
    for (int x = xlo; x <= xhi; x++) {
      int x0, x1;
      float lx;
      bilin_src(x, sw, Wi, x0, x1, lx);
      const float wx = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
      if (wx == 0.f) continue;
      float v[8];
      load8(dout + (((long)b * Ho + y) * Wo + x) * C + c, v);
      const float w = wy * wx;
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] = fmaf(w, v[i], acc[i]);
    }
  }
  store8(dz + pix * C + c, acc);
}
CMX_API int cmx_upsample_bwd(const void* dout, int Ho, int Wo, void* dz, int Hi, int Wi, int B, int C, void* stream) {
  CMX_REQUIRE(C % 8 == 0, "upsample_bwd: C %% 8");
  const long total = (long)B * Hi * Wi * (C >> 3);
  if (total == 0) return 0;
  upsample_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)dout, Ho, Wo, (bf16*)dz, Hi, Wi, B, C,
                                                                         (float)Hi / (float)Ho, (float)Wi / (float)Wo);
  LAUNCH_DONE("upsample_bwd");
}

// ---- adjoint of up to three bilinear sources in ONE pass over dout (row-separable form) -----------------------
// CTA = (destination k, low-res row yi, sample b, 64-channel chunk).  Phase 1 folds the <= 2*ceil(Ho/Hi)+2 output rows
// whose vertical stencil touches yi into a [Wo][64] fp32 row in shared memory (coalesced 16-byte loads, all threads,
// the row list is CTA-uniform so there is no divergence); phase 2 folds that row horizontally into the Wi low-res
// pixels.  blockIdx.x enumerates the rows of ALL destinations that belong to one band of the coarsest destination, so
// CTAs that read the same dout rows are adjacent in launch order and dout comes from DRAM about once instead of 3 x 4.
struct UpDst {
  bf16* dz;
  int H, W;
  float sh, sw;
  int rows_per_band;
  int maxc;  // upper bound of the horizontal stencil support in output pixels (2 * ceil(Wo / W) + 4)
};
constexpr int UPB_CC = 64;     // channels per CTA
constexpr int UPB_MAXROWS = 64;
__global__ void __launch_bounds__(256) upsample_bwd_multi_kernel(const bf16* __restrict__ dout, int Ho, int Wo, int C, UpDst d0,
                                                                 UpDst d1, UpDst d2, int nchunk) {
  pdl_trigger();
  extern __shared__ float s_t[];  // [Wo][UPB_CC] row accumulator, then the horizontal weight table
  __shared__ int s_y[UPB_MAXROWS];
  __shared__ float s_w[UPB_MAXROWS];
  __shared__ int s_n;
  int job = blockIdx.x;
  UpDst d = d0;
  if (job >= d.rows_per_band) {
    job -= d.rows_per_band;
    d = d1;
    if (job >= d.rows_per_band) {
      job -= d.rows_per_band;
      d = d2;
    }
  }
  const int yi = blockIdx.y * d.rows_per_band + job;
  if (yi >= d.H) return;
  const int b = blockIdx.z / nchunk;
  const int c0 = (blockIdx.z % nchunk) * UPB_CC;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const float rh = 1.f / d.sh;
    int ylo = (int)floorf((yi - 0.5f) * rh - 0.5f) - 1, yhi = (int)ceilf((yi + 1.5f) * rh - 0.5f) + 1;
    if (ylo < 0) ylo = 0;
    if (yhi > Ho - 1) yhi = Ho - 1;
    int n = 0;
    for (int y = ylo; y <= yhi; y++) {
      int y0, y1;
      float ly;
      bilin_src(y, d.sh, d.H, y0, y1, ly);
      const float wy = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
      if (wy != 0.f && n < UPB_MAXROWS) { s_y[n] = y; s_w[n] = wy; n++; }
    }
    s_n = n;
  }
  // horizontal weights, once per CTA instead of once per (pixel, channel pair): s_wx[xi][j] = weight of output column
  // s_x0[xi] + j in low-res column xi
  float* s_wx = s_t + Wo * UPB_CC;
  int* s_x0 = reinterpret_cast<int*>(s_wx + d.W * d.maxc);
  {
    const float rw = 1.f / d.sw;
    for (int i = tid; i < d.W * d.maxc; i += 256) {
      const int xi = i / d.maxc, j = i - xi * d.maxc;
      int xlo = (int)floorf((xi - 0.5f) * rw - 0.5f) - 1;
      if (xlo < 0) xlo = 0;
      const int x = xlo + j;
      float wx = 0.f;
      if (x < Wo) {
        int x0, x1;
        float lx;
        bilin_src(x, d.sw, d.W, x0, x1, lx);
        wx = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
      }
      s_wx[i] = wx;
      if (j == 0) s_x0[xi] = xlo;
    }
  }
  __syncthreads();
  const int n = s_n;
  // phase 1: t[x][c] = sum_rows wy * dout[b, y, x, c0 + c]
  const bf16* base = dout + (long)b * Ho * Wo * C + c0;
  for (int item = tid; item < Wo * (UPB_CC / 8); item += 256) {
    const int g = item & (UPB_CC / 8 - 1), x = item / (UPB_CC / 8);
    const bf16* px = base + (long)x * C + g * 8;
    float2 acc[4];
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i] = make_float2(0.f, 0.f);
    int r = 0;
    for (; r + 3 < n; r += 4) {  // four independent 16-byte loads in flight, packed fp32x2 FMAs
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; q++) u[q] = *reinterpret_cast<const uint4*>(px + (long)s_y[r + q] * Wo * C);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float w = s_w[r + q];
        const float2 w2 = make_float2(w, w);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u[q]);
#pragma unroll
        for (int i = 0; i < 4; i++) ffma2(acc[i], w2, __bfloat1622float2(h[i]));
      }
    }
    for (; r < n; r++) {
      const uint4 u = *reinterpret_cast<const uint4*>(px + (long)s_y[r] * Wo * C);
      const float w = s_w[r];
      const float2 w2 = make_float2(w, w);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; i++) ffma2(acc[i], w2, __bfloat1622float2(h[i]));
    }
    float4* o = reinterpret_cast<float4*>(s_t + x * UPB_CC + g * 8);
    o[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
    o[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
  }
  __syncthreads();
  // phase 2: dz[b, yi, xi, c0 + c] = sum_x wx * t[x][c]   (thread = channel pair of one low-res pixel; a warp shares xi,
  // so the weight reads are shared-memory broadcasts)
  bf16* orow = d.dz + ((long)b * d.H + yi) * d.W * C + c0;
  for (int item = tid; item < d.W * (UPB_CC / 2); item += 256) {
    const int cp = item & (UPB_CC / 2 - 1), xi = item / (UPB_CC / 2);
    const int xlo = s_x0[xi];
    int cnt = Wo - xlo;
    if (cnt > d.maxc) cnt = d.maxc;
    const float* wrow = s_wx + xi * d.maxc;
    const float* trow = s_t + xlo * UPB_CC + cp * 2;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < cnt; j++) {
      const float wx = wrow[j];
      const float2 tv = *reinterpret_cast<const float2*>(trow + j * UPB_CC);
      a0 = fmaf(wx, tv.x, a0);
      a1 = fmaf(wx, tv.y, a1);
    }
    *reinterpret_cast<__nv_bfloat162*>(orow + (long)xi * C + cp * 2) = __floats2bfloat162_rn(a0, a1);
  }
}
CMX_API int cmx_upsample_bwd_multi(const void* dout, int Ho, int Wo, void* dz1, int H1, int W1, void* dz2, int H2, int W2,
                                   void* dz3, int H3, int W3, int B, int C, void* stream) {
  CMX_REQUIRE(C % UPB_CC == 0, "upsample_bwd_multi: C %% 64");
  CMX_REQUIRE(dz1 && H1 > 0 && W1 > 0, "upsample_bwd_multi: at least one destination");
  if (B == 0) return 0;
  void* dz[3] = {dz1, dz2, dz3};
  int Hs[3] = {H1, H2, H3}, Ws[3] = {W1, W2, W3};
  int nd = 0, bands = 1 << 30;
  for (int k = 0; k < 3; k++)
    if (dz[k]) {
      CMX_REQUIRE(nd == k, "upsample_bwd_multi: destinations must be packed from the first slot");
      CMX_REQUIRE(Hs[k] > 0 && Ws[k] > 0 && Hs[k] <= Ho && Ws[k] <= Wo, "upsample_bwd_multi: destination %d larger than the source", k);
      CMX_REQUIRE(2 * cdiv(Ho, Hs[k]) + 4 <= UPB_MAXROWS, "upsample_bwd_multi: scale factor too large");
      nd++;
      if (Hs[k] < bands) bands = Hs[k];
    }
  UpDst d[3];
  int jobs = 0;
  size_t tab_bytes = 0;
  for (int k = 0; k < 3; k++) {
    if (k < nd) {
      d[k] = UpDst{(bf16*)dz[k], Hs[k], Ws[k], (float)Hs[k] / (float)Ho, (float)Ws[k] / (float)Wo, cdiv(Hs[k], bands),
                   2 * cdiv(Wo, Ws[k]) + 4};
      jobs += d[k].rows_per_band;
      const size_t tab = (size_t)Ws[k] * (d[k].maxc + 1) * sizeof(float);
      if (tab > tab_bytes) tab_bytes = tab;
    } else {
      d[k] = UpDst{nullptr, 0, 0, 1.f, 1.f, 1 << 30, 0};
    }
  }
  const int nchunk = C / UPB_CC;
  const size_t smem = (size_t)Wo * UPB_CC * sizeof(float) + tab_bytes;
  CMX_REQUIRE(smem <= 200 * 1024, "upsample_bwd_multi: Wo=%d too wide", Wo);
  CMX_REQUIRE((long)B * nchunk <= 65535 && bands <= 65535, "upsample_bwd_multi: grid too large");
  if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_bwd_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(jobs, bands, B * nchunk);
  upsample_bwd_multi_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)dout, Ho, Wo, C, d[0], d[1], d[2], nchunk);
  LAUNCH_DONE("upsample_bwd_multi");
}

// ---- fused bilinear upsample + cross entropy (+ gradient scatter) ----------------------------------------
// CTA = 32 x 8 threads covering a 32 x 8 full-resolution tile.  The low-res window the tile touches is
// staged in shared memory (logits in, gradient accumulator out); MAXC classes.
constexpr int CE_MAXC_SMALL = 16, CE_MAXC_LARGE = 64;  // register-array bound on the class count (two instantiations)
constexpr int CE_TW = 32, CE_TH = 8;
constexpr float CE_FIX = 2097152.f;  // 2^21: fixed-point scale of the per-CTA gradient partials (plain CE)

// FOCAL: loss = w_ce * CE + w_focal * FocalLoss, where FocalLoss is the reference's all-classes form (utils/loss_opr.py:
// 157-196: pt_k = p_k for the target class and 1 - p_k otherwise, -alpha_k (1 - pt_k)^gamma log(pt_k + 1e-8) summed over the
// classes; alpha_k = alpha / 1 - alpha), both averaged over the valid pixels (train.py:70-93: 'FocalLoss', 'CE_Focal').
struct FocalArgs { float w_ce, w_focal, gamma, alpha; };
template <bool FOCAL, int CE_MAXC>
__global__ void __launch_bounds__(256) ce_upsampled_kernel(const float* __restrict__ logits, long ld, const int64_t* __restrict__ label,
                                                           int ignore_index, double* __restrict__ acc, float* __restrict__ dlogits,
                                                           int h, int w, int H, int W, int ncls, float sh, float sw, int cap,
                                                           FocalArgs fa) {
  pdl_trigger();
  extern __shared__ float s_dyn[];  // [cap] low-res logits window, [cap] gradient accumulator
  float* s_l = s_dyn;
  float* s_g = s_dyn + cap;
  __shared__ float s_red[8];
  __shared__ int s_cnt[8];
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * CE_TW, Y0 = blockIdx.y * CE_TH;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  // low-res window [ly0, ly1] x [lx0, lx1] touched by this tile
  int ly0, t1, lx0;
  float tf;
  bilin_src(Y0, sh, h, ly0, t1, tf);
  bilin_src(X0, sw, w, lx0, t1, tf);
  int ylast = Y0 + CE_TH - 1, xlast = X0 + CE_TW - 1;
  if (ylast > H - 1) ylast = H - 1;
  if (xlast > W - 1) xlast = W - 1;
  int ly1, lx1, t0;
  bilin_src(ylast, sh, h, t0, ly1, tf);
  bilin_src(xlast, sw, w, t0, lx1, tf);
  const int nh = ly1 - ly0 + 1, nw = lx1 - lx0 + 1;
  if (nh * nw * ncls > cap) __trap();  // host sizing bug
  for (int i = tid; i < nh * nw * ncls; i += 256) {
    const int k = i % ncls;
    const int xx = (i / ncls) % nw;
    const int yy = i / (ncls * nw);
    s_l[i] = logits[(((long)b * h + ly0 + yy) * w + lx0 + xx) * ld + k];
    s_g[i] = 0.f;
  }
  __syncthreads();
  const int x = X0 + threadIdx.x, y = Y0 + threadIdx.y;
  float loss = 0.f;
  int valid = 0;
  if (x < W && y < H) {
    const long lab = label[((long)b * H + y) * W + x];
    if (lab != ignore_index) {
      // a target outside [0, ncls) that is not the ignore index is a data error: torch's nll_loss raises a device-side
      // assert for it; silently treating it as "no target" would train on garbage (FocalLoss clamps instead, like the reference)
      if (!FOCAL && (lab < 0 || lab >= ncls)) __trap();
      int y0, y1, x0, x1;
      float ly, lx;
      bilin_src(y, sh, h, y0, y1, ly);
      bilin_src(x, sw, w, x0, x1, lx);
      const int o00 = ((y0 - ly0) * nw + (x0 - lx0)) * ncls, o01 = ((y0 - ly0) * nw + (x1 - lx0)) * ncls;
      const int o10 = ((y1 - ly0) * nw + (x0 - lx0)) * ncls, o11 = ((y1 - ly0) * nw + (x1 - lx0)) * ncls;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      float v[CE_MAXC];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < CE_MAXC; k++) {
        if (k < ncls) {
          v[k] = w00 * s_l[o00 + k] + w01 * s_l[o01 + k] + w10 * s_l[o10 + k] + w11 * s_l[o11 + k];
          mx = fmaxf(mx, v[k]);
        }
      }
      float sum = 0.f, picked = 0.f;
#pragma unroll
      for (int k = 0; k < CE_MAXC; k++) {
        if (k < ncls) {
          if (k == (int)lab) picked = v[k];
          v[k] = expf(v[k] - mx);
          sum += v[k];
        }
      }
      loss = logf(sum) + mx - picked;
      valid = 1;
      const float inv = 1.f / sum;
      float hk[CE_MAXC], hs = 0.f;   // FOCAL: d focal / d p_k and sum_k h_k p_k
      if (FOCAL) {
        // out-of-range labels are clamped like the reference (loss_opr.py:171) - CE has no such case (it would raise)
        const int t = lab < 0 ? 0 : (lab >= ncls ? ncls - 1 : (int)lab);
        float fl = 0.f;
#pragma unroll
        for (int k = 0; k < CE_MAXC; k++) {
          hk[k] = 0.f;
          if (k < ncls) {
            const float p = v[k] * inv;
            if (k == t) {
              const float om = fmaxf(1.f - p, 0.f), lg = logf(p + 1e-8f);
              const float pw1 = om > 0.f ? __powf(om, fa.gamma - 1.f) : (fa.gamma == 1.f ? 1.f : 0.f);
              fl -= fa.alpha * pw1 * om * lg;
              hk[k] = fa.alpha * (fa.gamma * pw1 * lg - pw1 * om / (p + 1e-8f));
            } else {
              const float q = fmaxf(1.f - p, 0.f), lg = logf(q + 1e-8f);
              const float pw1 = p > 0.f ? __powf(p, fa.gamma - 1.f) : (fa.gamma == 1.f ? 1.f : 0.f);
              fl -= (1.f - fa.alpha) * pw1 * p * lg;
              hk[k] = (1.f - fa.alpha) * (-fa.gamma * pw1 * lg + pw1 * p / (q + 1e-8f));
            }
            hs += hk[k] * p;
          }
        }
        loss = fa.w_ce * loss + fa.w_focal * fl;
      }
      if (dlogits) {
#pragma unroll
        for (int k = 0; k < CE_MAXC; k++) {
          if (k < ncls) {
            float g = v[k] * inv - (k == (int)lab ? 1.f : 0.f);
            if (FOCAL) g = fa.w_ce * g + fa.w_focal * (v[k] * inv) * (hk[k] - hs);
            if (FOCAL) {
              atomicAdd(&s_g[o00 + k], w00 * g);
              atomicAdd(&s_g[o01 + k], w01 * g);
              atomicAdd(&s_g[o10 + k], w10 * g);
              atomicAdd(&s_g[o11 + k], w11 * g);
            } else {
              // plain CE: |w * g| <= 1 and a CTA has 256 pixels, so the per-CTA partial sums fit a 2^-21 fixed-point
              // int32: native shared-memory integer atomics (ATOMS.ADD) instead of 36 compare-and-swap loops per pixel
              // on addresses that 16 neighbouring pixels share, and an order-independent (deterministic) CTA partial
              int* s_gi = reinterpret_cast<int*>(s_g);
              atomicAdd(&s_gi[o00 + k], __float2int_rn(w00 * g * CE_FIX));
              atomicAdd(&s_gi[o01 + k], __float2int_rn(w01 * g * CE_FIX));
              atomicAdd(&s_gi[o10 + k], __float2int_rn(w10 * g * CE_FIX));
              atomicAdd(&s_gi[o11 + k], __float2int_rn(w11 * g * CE_FIX));
            }
          }
        }
      }
    }
  }
  // block reduction of loss / count
  loss = warp_sum(loss);
  valid = __reduce_add_sync(0xffffffffu, valid);
  if (threadIdx.x == 0) { s_red[threadIdx.y] = loss; s_cnt[threadIdx.y] = valid; }
  __syncthreads();
  if (tid == 0) {
    double l = 0.;
    int n = 0;
    for (int i = 0; i < 8; i++) { l += (double)s_red[i]; n += s_cnt[i]; }
    if (n) { atomicAdd(acc, l); atomicAdd(acc + 1, (double)n); }
  }
  if (dlogits) {
    for (int i = tid; i < nh * nw * ncls; i += 256) {
      const int k = i % ncls;
      const int xx = (i / ncls) % nw;
      const int yy = i / (ncls * nw);
      const float g = FOCAL ? s_g[i] : (float)reinterpret_cast<const int*>(s_g)[i] * (1.f / CE_FIX);
      if (g != 0.f) atomicAdd(dlogits + (((long)b * h + ly0 + yy) * w + lx0 + xx) * ld + k, g);
    }
  }
}
static int ce_launch(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc, float* dlogits, int B,
                     int h, int w, int H, int W, int ncls, const FocalArgs* focal, void* stream) {
  CMX_REQUIRE(ncls >= 1 && ncls <= CE_MAXC_LARGE, "ce: ncls=%d > %d unsupported", ncls, CE_MAXC_LARGE);
  CMX_REQUIRE(ld >= ncls, "ce: row stride %ld < ncls %d", (long)ld, ncls);
  CMX_REQUIRE(H >= h && W >= w, "ce: only upsampling supported");
  if (B == 0) return 0;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int nh_max = (int)((CE_TH - 1) * sh) + 3, nw_max = (int)((CE_TW - 1) * sw) + 3;
  const int cap = nh_max * nw_max * ncls;
  const size_t smem = (size_t)cap * 2 * sizeof(float);
  CMX_REQUIRE(smem <= 200 * 1024, "ce: low-res window too large for shared memory");
  dim3 grid(cdiv(W, CE_TW), cdiv(H, CE_TH), B), block(32, 8);
  const FocalArgs fa = focal ? *focal : FocalArgs{1.f, 0.f, 0.f, 0.f};
#define CE_LAUNCH(F, MC) \
  if (smem > 48 * 1024) cudaFuncSetAttribute(ce_upsampled_kernel<F, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  ce_upsampled_kernel<F, MC><<<grid, block, smem, (cudaStream_t)stream>>>(logits, (long)ld, label, ignore_index, acc, dlogits, h, w, H, W, ncls, sh, sw, cap, fa)
  if (ncls <= CE_MAXC_SMALL) { if (focal) { CE_LAUNCH(true, CE_MAXC_SMALL); } else { CE_LAUNCH(false, CE_MAXC_SMALL); } }
  else { if (focal) { CE_LAUNCH(true, CE_MAXC_LARGE); } else { CE_LAUNCH(false, CE_MAXC_LARGE); } }   // e.g. the 40 NYUDv2 classes (config.py default)
#undef CE_LAUNCH
  LAUNCH_DONE("ce_upsampled_fwd_bwd");
}
CMX_API int cmx_ce_upsampled_fwd_bwd(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc,
                                     float* dlogits, int B, int h, int w, int H, int W, int ncls, void* stream) {
  return ce_launch(logits, ld, label, ignore_index, acc, dlogits, B, h, w, H, W, ncls, nullptr, stream);
}
CMX_API int cmx_ce_focal_upsampled_fwd_bwd(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc,
                                           float* dlogits, int B, int h, int w, int H, int W, int ncls, float w_ce, float w_focal,
                                           float gamma, float alpha, void* stream) {
  CMX_REQUIRE(gamma >= 0.f && alpha >= 0.f && alpha <= 1.f, "ce_focal: gamma=%g alpha=%g out of range", gamma, alpha);
  const FocalArgs fa{w_ce, w_focal, gamma, alpha};
  return ce_launch(logits, ld, label, ignore_index, acc, dlogits, B, h, w, H, W, ncls, &fa, stream);
}

// ---- DiceCELoss (utils/loss_opr.py:103-156): alpha * Dice + (1 - alpha) * CE on the bilinearly upsampled logits -----------
// Dice couples all pixels of a sample through per-(sample, class) sums, so it takes two passes over the pixels (the
// upsampled logits are recomputed from the shared-memory low-res window in both; nothing full-resolution is stored):
//   PASS 0  per (b, k):  Sp = sum_valid p_k,  Si = sum_valid p_k [k == t],  St = sum_valid [k == t]   (t = label clamped to
//           [0, ncls) as the reference does) and the CE sum / valid count;
//   finalize (one CTA): dice_bk = (2 Si + s) / (Sp + St + s), loss = alpha (1 - mean dice) + (1 - alpha) CE, and the
//           coefficients of d loss / d p_k = c1_bk [k == t] + c2_bk;
//   PASS 1  d loss / d z_j = p_j (g_j - sum_k g_k p_k) + (1 - alpha) / n_valid (p_j - [j == t]), scattered through the
//           bilinear weights to the low-res logits exactly like the CE kernel does.
template <int PASS, int MAXC>
__global__ void __launch_bounds__(256) dice_ce_kernel(const float* __restrict__ logits, long ld, const int64_t* __restrict__ label,
                                                      int ignore_index, double* __restrict__ acc, double* __restrict__ dstats,
                                                      const float* __restrict__ coef, float* __restrict__ dlogits,
                                                      int h, int w, int H, int W, int ncls, float sh, float sw, int cap) {
  pdl_trigger();
  extern __shared__ float s_dyn[];  // [cap] low-res logits window, [cap] gradient accumulator, [3 * ncls] class sums
  float* s_l = s_dyn;
  float* s_g = s_dyn + cap;
  float* s_c = s_dyn + 2 * cap;
  __shared__ float s_red[8];
  __shared__ int s_cnt[8];
  const int b = blockIdx.z;
  const int X0 = blockIdx.x * CE_TW, Y0 = blockIdx.y * CE_TH;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  int ly0, t1, lx0;
  float tf;
  bilin_src(Y0, sh, h, ly0, t1, tf);
  bilin_src(X0, sw, w, lx0, t1, tf);
  int ylast = Y0 + CE_TH - 1, xlast = X0 + CE_TW - 1;
  if (ylast > H - 1) ylast = H - 1;
  if (xlast > W - 1) xlast = W - 1;
  int ly1, lx1, t0;
  bilin_src(ylast, sh, h, t0, ly1, tf);
  bilin_src(xlast, sw, w, t0, lx1, tf);
  const int nh = ly1 - ly0 + 1, nw = lx1 - lx0 + 1;
  if (nh * nw * ncls > cap) __trap();
  for (int i = tid; i < nh * nw * ncls; i += 256) {
    const int k = i % ncls;
    const int xx = (i / ncls) % nw;
    const int yy = i / (ncls * nw);
    s_l[i] = logits[(((long)b * h + ly0 + yy) * w + lx0 + xx) * ld + k];
    s_g[i] = 0.f;
  }
  for (int i = tid; i < 3 * ncls; i += 256) s_c[i] = 0.f;
  __syncthreads();
  const int x = X0 + threadIdx.x, y = Y0 + threadIdx.y;
  float loss = 0.f;
  int valid = 0;
  if (x < W && y < H) {
    const long lab = label[((long)b * H + y) * W + x];
    if (lab != ignore_index) {
      if (lab < 0 || lab >= ncls) __trap();   // the CE term of the reference raises for such a target
      const int t = (int)lab;
      int y0, y1, x0, x1;
      float ly, lx;
      bilin_src(y, sh, h, y0, y1, ly);
      bilin_src(x, sw, w, x0, x1, lx);
      const int o00 = ((y0 - ly0) * nw + (x0 - lx0)) * ncls, o01 = ((y0 - ly0) * nw + (x1 - lx0)) * ncls;
      const int o10 = ((y1 - ly0) * nw + (x0 - lx0)) * ncls, o11 = ((y1 - ly0) * nw + (x1 - lx0)) * ncls;
      const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
      float v[MAXC];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < MAXC; k++) {
        if (k < ncls) {
          v[k] = w00 * s_l[o00 + k] + w01 * s_l[o01 + k] + w10 * s_l[o10 + k] + w11 * s_l[o11 + k];
          mx = fmaxf(mx, v[k]);
        }
      }
      float sum = 0.f, picked = 0.f;
#pragma unroll
      for (int k = 0; k < MAXC; k++) {
        if (k < ncls) {
          if (k == t) picked = v[k];
          v[k] = expf(v[k] - mx);
          sum += v[k];
        }
      }
      const float inv = 1.f / sum;
      valid = 1;
      if (PASS == 0) {
        loss = logf(sum) + mx - picked;
#pragma unroll
        for (int k = 0; k < MAXC; k++) {
          if (k < ncls) {
            const float p = v[k] * inv;
            atomicAdd(&s_c[k], p);
            if (k == t) { atomicAdd(&s_c[ncls + k], p); atomicAdd(&s_c[2 * ncls + k], 1.f); }
          }
        }
      } else {
        const float* c1 = coef + (long)b * 2 * ncls;
        const float* c2 = c1 + ncls;
        const float w_ce = coef[(long)gridDim.z * 2 * ncls];   // (1 - alpha) / n_valid, written by the finalize kernel
        float gs = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; k++)
          if (k < ncls) gs += ((k == t ? c1[k] : 0.f) + c2[k]) * v[k] * inv;
#pragma unroll
        for (int k = 0; k < MAXC; k++) {
          if (k < ncls) {
            const float p = v[k] * inv;
            const float g = p * ((k == t ? c1[k] : 0.f) + c2[k] - gs) + w_ce * (p - (k == t ? 1.f : 0.f));
            atomicAdd(&s_g[o00 + k], w00 * g);
            atomicAdd(&s_g[o01 + k], w01 * g);
            atomicAdd(&s_g[o10 + k], w10 * g);
            atomicAdd(&s_g[o11 + k], w11 * g);
          }
        }
      }
    }
  }
  if (PASS == 0) {
    loss = warp_sum(loss);
    valid = __reduce_add_sync(0xffffffffu, valid);
    if (threadIdx.x == 0) { s_red[threadIdx.y] = loss; s_cnt[threadIdx.y] = valid; }
    __syncthreads();
    if (tid == 0) {
      double l = 0.;
      int n = 0;
      for (int i = 0; i < 8; i++) { l += (double)s_red[i]; n += s_cnt[i]; }
      if (n) { atomicAdd(acc, l); atomicAdd(acc + 1, (double)n); }
    }
    for (int i = tid; i < 3 * ncls; i += 256)
      if (s_c[i] != 0.f) atomicAdd(dstats + (long)b * 3 * ncls + i, (double)s_c[i]);
  } else {
    __syncthreads();
    for (int i = tid; i < nh * nw * ncls; i += 256) {
      const int k = i % ncls;
      const int xx = (i / ncls) % nw;
      const int yy = i / (ncls * nw);
      if (s_g[i] != 0.f) atomicAdd(dlogits + (((long)b * h + ly0 + yy) * w + lx0 + xx) * ld + k, s_g[i]);
    }
  }
}
static int dice_launch(int pass, const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc, double* dstats,
                       const float* coef, float* dlogits, int B, int h, int w, int H, int W, int ncls, void* stream) {
  CMX_REQUIRE(ncls >= 1 && ncls <= CE_MAXC_LARGE, "dice_ce: ncls=%d > %d unsupported", ncls, CE_MAXC_LARGE);
  CMX_REQUIRE(ld >= ncls, "dice_ce: row stride %ld < ncls %d", (long)ld, ncls);
  CMX_REQUIRE(H >= h && W >= w, "dice_ce: only upsampling supported");
  if (B == 0) return 0;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int nh_max = (int)((CE_TH - 1) * sh) + 3, nw_max = (int)((CE_TW - 1) * sw) + 3;
  const int cap = nh_max * nw_max * ncls;
  const size_t smem = ((size_t)cap * 2 + 3 * ncls) * sizeof(float);
  CMX_REQUIRE(smem <= 200 * 1024, "dice_ce: low-res window too large for shared memory");
  dim3 grid(cdiv(W, CE_TW), cdiv(H, CE_TH), B), block(32, 8);
#define DICE_LAUNCH(P, MC) \
  if (smem > 48 * 1024) cudaFuncSetAttribute(dice_ce_kernel<P, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  dice_ce_kernel<P, MC><<<grid, block, smem, (cudaStream_t)stream>>>(logits, (long)ld, label, ignore_index, acc, dstats, coef, dlogits, h, w, H, W, ncls, sh, sw, cap)
  if (ncls <= CE_MAXC_SMALL) { if (pass == 0) { DICE_LAUNCH(0, CE_MAXC_SMALL); } else { DICE_LAUNCH(1, CE_MAXC_SMALL); } }
  else { if (pass == 0) { DICE_LAUNCH(0, CE_MAXC_LARGE); } else { DICE_LAUNCH(1, CE_MAXC_LARGE); } }
#undef DICE_LAUNCH
  LAUNCH_DONE("dice_ce");
}
CMX_API int cmx_dice_ce_stats(const float* logits, int64_t ld, const int64_t* label, int ignore_index, double* acc, double* dstats,
                              int B, int h, int w, int H, int W, int ncls, void* stream) {
  return dice_launch(0, logits, ld, label, ignore_index, acc, dstats, nullptr, nullptr, B, h, w, H, W, ncls, stream);
}
// one CTA: loss and the gradient coefficients from the pass-0 sums.  coef: float[B][2][ncls] = (c1, c2) followed by one float
// = (1 - alpha) / n_valid (read by pass 1)
__global__ void dice_ce_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ dstats, int B, int ncls, float alpha,
                                        float smooth, float* __restrict__ loss, float* __restrict__ coef) {
  pdl_trigger();
  __shared__ double s_sum[256];
  double part = 0.;
  const int n = B * ncls;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / ncls, k = i % ncls;
    const double* st = dstats + (long)b * 3 * ncls;
    const double U = st[k] + st[2 * ncls + k] + (double)smooth, num = 2. * st[ncls + k] + (double)smooth;
    part += num / U;
    if (coef) {
      coef[(long)b * 2 * ncls + k] = (float)(-(double)alpha / n * 2. / U);
      coef[(long)b * 2 * ncls + ncls + k] = (float)((double)alpha / n * num / (U * U));
    }
  }
  s_sum[threadIdx.x] = part;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double dice_loss = 1. - s_sum[0] / n;
    const double ce = acc[0] / acc[1];          // all-ignored batch -> NaN like the reference's CE term
    if (loss) *loss = (float)((double)alpha * dice_loss + (1. - (double)alpha) * ce);
    if (coef) coef[(long)B * 2 * ncls] = (float)((1. - (double)alpha) / acc[1]);
  }
}
CMX_API int cmx_dice_ce_finalize(const double* acc, const double* dstats, int B, int ncls, float alpha, float smooth, float* loss,
                                 float* coef, void* stream) {
  dice_ce_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(acc, dstats, B, ncls, alpha, smooth, loss, coef);
  LAUNCH_DONE("dice_ce_finalize");
}
// pass 1: dlogits (fp32, zero-initialised by the caller) += d loss / d logits; coef as written by cmx_dice_ce_finalize
// (float[B*2*ncls + 1], the last element = (1 - alpha) / n_valid); the result needs no further scaling
CMX_API int cmx_dice_ce_grad(const float* logits, int64_t ld, const int64_t* label, int ignore_index, const float* coef,
                             float* dlogits, int B, int h, int w, int H, int W, int ncls, void* stream) {
  return dice_launch(1, logits, ld, label, ignore_index, nullptr, nullptr, coef, dlogits, B, h, w, H, W, ncls, stream);
}
template <typename TO>
__global__ void __launch_bounds__(256) ce_finalize_kernel(const double* __restrict__ acc, float* __restrict__ loss,
                                                          const float* __restrict__ dlogits, const float* __restrict__ gscale,
                                                          TO* __restrict__ out, long n) {
  pdl_trigger();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && loss) *loss = (float)(acc[0] / acc[1]);  // all-ignored batch -> 0/0 = NaN, like torch
  if (i < n && out) {
    const float s = (gscale ? *gscale : 1.f) / (float)acc[1];
    st1(out + i, dlogits[i] * s);
  }
}
CMX_API int cmx_ce_finalize(const double* acc, float* loss, const float* dlogits, const float* gscale, void* dlogits_out,
                            int out_dtype, int64_t n, void* stream) {
  const long nn = dlogits_out ? n : 1;
  if (out_dtype == CMX_F32) ce_finalize_kernel<float><<<cdiv(nn, 256), 256, 0, (cudaStream_t)stream>>>(acc, loss, dlogits, gscale, (float*)dlogits_out, n);
  else ce_finalize_kernel<bf16><<<cdiv(nn, 256), 256, 0, (cudaStream_t)stream>>>(acc, loss, dlogits, gscale, (bf16*)dlogits_out, n);
  LAUNCH_DONE("ce_finalize");
}

// ---- eval: low-res channels-last logits -> full-res NCHW ------------------------------------------------
__global__ void __launch_bounds__(256) logits_upsample_nchw_kernel(const float* __restrict__ logits, long ld, float* __restrict__ out, int B, int h,
                                                                   int w, int H, int W, int ncls, float sh, float sw) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / ((long)W * H));
  int y0, y1, x0, x1;
  float ly, lx;
  bilin_src(y, sh, h, y0, y1, ly);
  bilin_src(x, sw, w, x0, x1, lx);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const float* p00 = logits + (((long)b * h + y0) * w + x0) * ld;
  const float* p01 = logits + (((long)b * h + y0) * w + x1) * ld;
  const float* p10 = logits + (((long)b * h + y1) * w + x0) * ld;
  const float* p11 = logits + (((long)b * h + y1) * w + x1) * ld;
  for (int k = 0; k < ncls; k++)
    out[(((long)b * ncls + k) * H + y) * W + x] = w00 * p00[k] + w01 * p01[k] + w10 * p10[k] + w11 * p11[k];
}
CMX_API int cmx_logits_upsample_nchw(const float* logits, int64_t ld, float* out, int B, int h, int w, int H, int W, int ncls,
                                     void* stream) {
  CMX_REQUIRE(ld >= ncls, "logits_upsample: row stride %ld < ncls %d", (long)ld, ncls);
  const long total = (long)B * H * W;
  if (total == 0) return 0;
  logits_upsample_nchw_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(logits, (long)ld, out, B, h, w, H, W, ncls,
                                                                                 (float)h / (float)H, (float)w / (float)W);
  LAUNCH_DONE("logits_upsample_nchw");
}

// ---- confusion matrix (integer, bit-exact) -----------------------------------------------------------------
constexpr int CF_MAXCL = 64;  // n_cl^2 <= 4096 shared-memory bins
template <typename TP, typename TG>
__global__ void __launch_bounds__(256) confusion_kernel(const TP* __restrict__ pred, const TG* __restrict__ gt, long n, int n_cl,
                                                        unsigned long long* __restrict__ hist, unsigned long long* __restrict__ stats) {
  pdl_trigger();
  __shared__ unsigned int sh[CF_MAXCL * CF_MAXCL];
  __shared__ unsigned int s_lab, s_cor;
  for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) { s_lab = 0; s_cor = 0; }
  __syncthreads();
  unsigned int lab = 0, cor = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long g = (long)gt[i];
    if (g >= 0 && g < n_cl) {
      const long p = (long)pred[i];
      lab++;
      if (p == g) cor++;
      // numpy bincount(n_cl*gt+pred, minlength=n_cl^2): predictions are class ids in [0, n_cl)
      if (p >= 0 && p < n_cl) atomicAdd(&sh[g * n_cl + p], 1u);
    }
  }
  lab = __reduce_add_sync(0xffffffffu, lab);
  cor = __reduce_add_sync(0xffffffffu, cor);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_lab, lab); atomicAdd(&s_cor, cor); }
  __syncthreads();
  for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
  if (threadIdx.x == 0) { atomicAdd(stats, (unsigned long long)s_lab); atomicAdd(stats + 1, (unsigned long long)s_cor); }
}
template <typename TP>
static void confusion_launch(const TP* pred, const void* gt, int gt_dtype, long n, int n_cl, int64_t* hist, int64_t* stats, cudaStream_t st) {
  int grid = cdiv(n, 256 * 8);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  auto* h = reinterpret_cast<unsigned long long*>(hist);
  auto* s = reinterpret_cast<unsigned long long*>(stats);
  if (gt_dtype == 0) confusion_kernel<TP, uint8_t><<<grid, 256, 0, st>>>(pred, (const uint8_t*)gt, n, n_cl, h, s);
  else if (gt_dtype == 1) confusion_kernel<TP, int32_t><<<grid, 256, 0, st>>>(pred, (const int32_t*)gt, n, n_cl, h, s);
  else confusion_kernel<TP, int64_t><<<grid, 256, 0, st>>>(pred, (const int64_t*)gt, n, n_cl, h, s);
}
CMX_API int cmx_confusion(const void* pred, int pred_dtype, const void* gt, int gt_dtype, int64_t n, int n_cl, int64_t* hist,
                          int64_t* stats, void* stream) {
  CMX_REQUIRE(n_cl >= 1 && n_cl <= CF_MAXCL, "confusion: n_cl=%d unsupported (max %d)", n_cl, CF_MAXCL);
  CMX_REQUIRE(pred_dtype >= 0 && pred_dtype <= 2 && gt_dtype >= 0 && gt_dtype <= 2, "confusion: bad dtype tag");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (pred_dtype == 0) confusion_launch<uint8_t>((const uint8_t*)pred, gt, gt_dtype, n, n_cl, hist, stats, st);
  else if (pred_dtype == 1) confusion_launch<int32_t>((const int32_t*)pred, gt, gt_dtype, n, n_cl, hist, stats, st);
  else confusion_launch<int64_t>((const int64_t*)pred, gt, gt_dtype, n, n_cl, hist, stats, st);
  LAUNCH_DONE("confusion");
}

// fused argmax (first maximum, like numpy/torch argmax) + confusion for one [ncls, npix] score map
template <typename TG, typename TS>
__global__ void __launch_bounds__(256) argmax_confusion_kernel(const TS* __restrict__ scores, const TG* __restrict__ gt, long npix,
                                                               int n_cl, uint8_t* __restrict__ pred_out,
                                                               unsigned long long* __restrict__ hist, unsigned long long* __restrict__ stats) {
  pdl_trigger();
  __shared__ unsigned int sh[CF_MAXCL * CF_MAXCL];
  __shared__ unsigned int s_lab, s_cor;
  for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x == 0) { s_lab = 0; s_cor = 0; }
  __syncthreads();
  unsigned int lab = 0, cor = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
    TS best = scores[i];
    int p = 0;
    for (int k = 1; k < n_cl; k++) {
      const TS v = scores[(long)k * npix + i];
      if (v > best) { best = v; p = k; }
    }
    if (pred_out) pred_out[i] = (uint8_t)p;
    if (gt) {
      const long g = (long)gt[i];
      if (g >= 0 && g < n_cl) {
        lab++;
        if (p == g) cor++;
        atomicAdd(&sh[g * n_cl + p], 1u);
      }
    }
  }
  lab = __reduce_add_sync(0xffffffffu, lab);
  cor = __reduce_add_sync(0xffffffffu, cor);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_lab, lab); atomicAdd(&s_cor, cor); }
  __syncthreads();
  if (gt) {
    for (int i = threadIdx.x; i < n_cl * n_cl; i += blockDim.x)
      if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
    if (threadIdx.x == 0) { atomicAdd(stats, (unsigned long long)s_lab); atomicAdd(stats + 1, (unsigned long long)s_cor); }
  }
}
CMX_API int cmx_argmax_confusion(const void* scores, int score_f64, const void* gt, int gt_dtype, int64_t npix, int n_cl,
                                 uint8_t* pred_out, int64_t* hist, int64_t* stats, void* stream) {
  CMX_REQUIRE(n_cl >= 1 && n_cl <= CF_MAXCL, "argmax_confusion: n_cl=%d unsupported", n_cl);
  if (npix == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = cdiv(npix, 256 * 4);
  if (grid > 148 * 4) grid = 148 * 4;
  auto* h = reinterpret_cast<unsigned long long*>(hist);
  auto* s = reinterpret_cast<unsigned long long*>(stats);
#define AC_L(TG, TS) argmax_confusion_kernel<TG, TS><<<grid, 256, 0, st>>>((const TS*)scores, (const TG*)gt, npix, n_cl, pred_out, h, s)
  if (score_f64) {   // the evaluator sums the per-scale score maps in float64 (evaluator.py:309, 320)
    if (gt_dtype == 0) AC_L(uint8_t, double); else if (gt_dtype == 1) AC_L(int32_t, double); else AC_L(int64_t, double);
  } else {
    if (gt_dtype == 0) AC_L(uint8_t, float); else if (gt_dtype == 1) AC_L(int32_t, float); else AC_L(int64_t, float);
  }
#undef AC_L
  LAUNCH_DONE("argmax_confusion");
}
