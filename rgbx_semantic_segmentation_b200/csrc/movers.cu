// Layout movers and small streaming kernels: im2col / col2im for the overlapped patch embeds and the
// spatial-reduction conv, conv-weight packing, casts, column sums (bias grads), ReLU backward,
// row softmax (self-attention) and dim -2 softmax (FFM cross-attention context).
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

// ---- stage-1 im2col from the NCHW fp32 image ------------------------------------------------------
// I = index type of the thread -> (row, column group) decomposition: unsigned 32-bit whenever the element count allows
// (64-bit divisions cost ~100 instructions each and these kernels move 16-32 bytes per thread)
// CIN / KK > 0: compile-time channel count and kernel size (the 7x7 stride-4 stage-1 patch embed over 3 channels is the only
// caller that matters: 147 runtime divisions per output pixel otherwise); 0 = runtime values
template <typename I, int CIN, int KK>
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const float* __restrict__ x, bf16* __restrict__ col, int B, int Cin_, int H,
                                                          int W, int k_, int s, int p, int Ho, int Wo, int kpad) {
  pdl_trigger();
  const int Cin = CIN > 0 ? CIN : Cin_, k = KK > 0 ? KK : k_;
  // one thread = 8 consecutive im2col columns of one output pixel (one 16-byte store)
  const int g8 = kpad >> 3;
  const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
  const I total = (I)B * Ho * Wo * g8;
  if (idx >= total) return;
  const int j0 = (int)(idx % (I)g8) * 8;
  const I row = idx / (I)g8;
  const int ox = (int)(row % (I)Wo);
  const int oy = (int)((row / (I)Wo) % (I)Ho);
  const int b = (int)(row / ((I)Wo * Ho));
  const float* xb = x + (long)b * Cin * H * W;
  const int iy0 = oy * s - p, ix0 = ox * s - p;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int j = j0 + i;
    v[i] = 0.f;
    if (j < k * k * Cin) {
      const int ci = j % Cin, tap = j / Cin;
      const int kh = tap / k, kw = tap % k;
      const int iy = iy0 + kh, ix = ix0 + kw;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[i] = __ldg(xb + ((long)ci * H + iy) * W + ix);
    }
  }
  store8(col + (long)row * kpad + j0, v);
}
CMX_API int cmx_im2col_nchw(const float* x, void* col, int B, int Cin, int H, int W, int k, int s, int p, int Ho, int Wo,
                            int kpad, void* stream) {
  CMX_REQUIRE(kpad >= k * k * Cin && kpad % 8 == 0, "im2col_nchw: kpad must be a multiple of 8 and >= k*k*Cin");
  const long total = (long)B * Ho * Wo * (kpad >> 3);
  if (total == 0) return 0;
  if (total < (1L << 31) && Cin == 3 && k == 7)
    im2col_nchw_kernel<unsigned, 3, 7><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)col, B, Cin, H, W, k, s, p, Ho, Wo, kpad);
  else if (total < (1L << 31))
    im2col_nchw_kernel<unsigned, 0, 0><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)col, B, Cin, H, W, k, s, p, Ho, Wo, kpad);
  else
    im2col_nchw_kernel<long, 0, 0><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)col, B, Cin, H, W, k, s, p, Ho, Wo, kpad);
  LAUNCH_DONE("im2col_nchw");
}

// ---- input pipeline fused into the stage-1 patch-embed load: raw uint8 HWC image -> im2col rows ------------
// The reference normalises on the host (utils/transforms.py:182-187: (v / 255 - mean) / std in float64, cast to fp32), replicates
// the grey thermal image to 3 channels (RGBXDataset.py:57-59; each then normalised with ITS channel's mean / std, dataloader.py:
// 106) and transposes HWC -> CHW (dataloader.py:109-110); here the uint8 image is resident on the device and the float64
// normalisation happens while the 7x7/4 patches are gathered.
// ch = 3: column (tap, c) = normalised value (0 in the conv's zero padding) - bit-identical to the host pipeline + im2col_nchw.
// ch = 1 (grey X): the three replicated channels are affine images of ONE value, x_c = v/255 / std_c - mean_c / std_c, so the
//   7x7x3 weights fold into 7x7x2: column (tap, 0) = v / 255, column (tap, 1) = 1 inside the image (both 0 in the padding); the
//   caller multiplies by  Wa = sum_c W_c / std_c  and  Wb = -sum_c W_c mean_c / std_c  (98 instead of 147 columns).
template <typename I>
__global__ void __launch_bounds__(256) im2col_u8_kernel(const uint8_t* __restrict__ x, bf16* __restrict__ col, int B, int ch, int H, int W,
                                                        int k, int s, int p, int Ho, int Wo, int kpad, double m0, double m1, double m2,
                                                        double d0, double d1, double d2) {
  pdl_trigger();
  const int g8 = kpad >> 3;
  const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
  const I total = (I)B * Ho * Wo * g8;
  if (idx >= total) return;
  const int j0 = (int)(idx % (I)g8) * 8;
  const I row = idx / (I)g8;
  const int ox = (int)(row % (I)Wo);
  const int oy = (int)((row / (I)Wo) % (I)Ho);
  const int b = (int)(row / ((I)Wo * Ho));
  const int per_tap = ch == 1 ? 2 : ch;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int j = j0 + i;
    v[i] = 0.f;
    if (j < k * k * per_tap) {
      const int ci = j % per_tap, tap = j / per_tap;
      const int kh = tap / k, kw = tap % k;
      const int iy = oy * s - p + kh, ix = ox * s - p + kw;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {   // the conv's zero padding pads the NORMALISED image: stays 0
        if (ch == 1) {
          v[i] = ci == 0 ? (float)((double)x[((long)b * H + iy) * W + ix] / 255.0) : 1.f;
        } else {
          const double pv = (double)x[(((long)b * H + iy) * W + ix) * ch + ci];
          const double mean = ci == 0 ? m0 : (ci == 1 ? m1 : m2), sd = ci == 0 ? d0 : (ci == 1 ? d1 : d2);
          v[i] = (float)((pv / 255.0 - mean) / sd);
        }
      }
    }
  }
  store8(col + (long)row * kpad + j0, v);
}
CMX_API int cmx_im2col_u8(const uint8_t* x, void* col, int B, int ch, int H, int W, int k, int s, int p, int Ho, int Wo, int kpad,
                          double mean0, double mean1, double mean2, double std0, double std1, double std2, void* stream) {
  CMX_REQUIRE(ch == 1 || ch == 3, "im2col_u8: ch=%d (1 or 3)", ch);
  CMX_REQUIRE(kpad >= k * k * (ch == 1 ? 2 : ch) && kpad % 8 == 0, "im2col_u8: kpad must be a multiple of 8 and >= the column count");
  const long total = (long)B * Ho * Wo * (kpad >> 3);
  if (total == 0) return 0;
  if (total < (1L << 31))
    im2col_u8_kernel<unsigned><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)col, B, ch, H, W, k, s, p, Ho, Wo, kpad, mean0,
                                                                                 mean1, mean2, std0, std1, std2);
  else
    im2col_u8_kernel<long><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)col, B, ch, H, W, k, s, p, Ho, Wo, kpad, mean0, mean1,
                                                                             mean2, std0, std1, std2);
  LAUNCH_DONE("im2col_u8");
}

// ---- NHWC bf16 im2col (8 channels per thread) ------------------------------------------------------
template <typename I>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ col, int B, int H,
                                                          int W, int C, int k, int s, int p, int Ho, int Wo) {
  pdl_trigger();
  const int c8 = C >> 3;
  const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
  const I total = (I)B * Ho * Wo * k * k * c8;
  if (idx >= total) return;
  const int cg = (int)(idx % (I)c8);
  const I t = idx / (I)c8;
  const int tap = (int)(t % (I)(k * k));
  const I row = t / (I)(k * k);
  const int kh = tap / k, kw = tap % k;
  const int ox = (int)(row % (I)Wo);
  const int oy = (int)((row / (I)Wo) % (I)Ho);
  const int b = (int)(row / ((I)Wo * Ho));
  const int iy = oy * s - p + kh, ix = ox * s - p + kw;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(x + ((long)(b * H + iy) * W + ix) * ldx + cg * 8);
  *reinterpret_cast<uint4*>(col + ((long)row * (k * k) + tap) * C + cg * 8) = v;
}
// 32-bit variant with precomputed divisions (the production path: every call of the model has < 2^31 vectors)
struct Im2colDivs { FastDiv c8, kk, k, Wo, Ho; };
__global__ void __launch_bounds__(256) im2col_nhwc_fast_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ col, int H, int W, int C,
                                                               int s, int p, unsigned total, Im2colDivs dv) {
  pdl_trigger();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  unsigned t, cg, row, tap, kh, kw, q, ox, b, oy;
  fdivmod(idx, dv.c8, t, cg);
  fdivmod(t, dv.kk, row, tap);
  fdivmod(tap, dv.k, kh, kw);
  fdivmod(row, dv.Wo, q, ox);
  fdivmod(q, dv.Ho, b, oy);
  const int iy = (int)oy * s - p + (int)kh, ix = (int)ox * s - p + (int)kw;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(x + ((long)((int)b * H + iy) * W + ix) * ldx + cg * 8);
  *reinterpret_cast<uint4*>(col + (long)t * C + cg * 8) = v;   // t = row * k*k + tap
}
CMX_API int cmx_im2col_nhwc(const void* x, int64_t ldx, void* col, int B, int H, int W, int C, int k, int s, int p, int Ho,
                            int Wo, void* stream) {
  CMX_REQUIRE(C % 8 == 0 && ldx % 8 == 0, "im2col_nhwc: C %% 8");
  const long total = (long)B * Ho * Wo * k * k * (C >> 3);
  if (total == 0) return 0;
  if (total < (1L << 31)) {
    Im2colDivs dv{make_fastdiv((unsigned)(C >> 3)), make_fastdiv((unsigned)(k * k)), make_fastdiv((unsigned)k), make_fastdiv((unsigned)Wo),
                  make_fastdiv((unsigned)Ho)};
    im2col_nhwc_fast_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)col, H, W, C, s, p, (unsigned)total, dv);
  }
  else
    im2col_nhwc_kernel<long><<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)col, B, H, W, C, k, s, p, Ho, Wo);
  LAUNCH_DONE("im2col_nhwc");
}

// ---- col2im (adjoint of im2col_nhwc), gather form -------------------------------------------------
template <typename TA, typename TO, typename I>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const bf16* __restrict__ dcol, const TA* __restrict__ add, long ldadd,
                                                          TO* __restrict__ dx, long lddx, int B, int H, int W, int C, int k, int s,
                                                          int p, int Ho, int Wo) {
  pdl_trigger();
  const int c8 = C >> 3;
  const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
  const I total = (I)B * H * W * c8;
  if (idx >= total) return;
  const int cg = (int)(idx % (I)c8);
  const I pix = idx / (I)c8;
  const int ix = (int)(pix % (I)W);
  const int iy = (int)((pix / (I)W) % (I)H);
  const int b = (int)(pix / ((I)W * H));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  if (add) load8(add + (long)pix * ldadd + cg * 8, acc);
  // only taps with (iy + p - kh) divisible by the stride contribute: kh = (iy + p) mod s, + s, ... (one tap for the
  // non-overlapping SR convolutions k = s = R instead of a scan over all R*R)
  for (int kh = (iy + p) % s; kh < k; kh += s) {
    const int ty = iy + p - kh;
    if (ty < 0) break;
    const int oy = ty / s;
    if (oy >= Ho) continue;
    for (int kw = (ix + p) % s; kw < k; kw += s) {
      const int tx = ix + p - kw;
      if (tx < 0) break;
      const int ox = tx / s;
      if (ox >= Wo) continue;
      float v[8];
      load8(dcol + ((((long)b * Ho + oy) * Wo + ox) * (k * k) + kh * k + kw) * C + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += v[i];
    }
  }
  store8(dx + (long)pix * lddx + cg * 8, acc);
}
struct Col2imDivs { FastDiv c8, W, H, s; };
template <typename TA, typename TO>
__global__ void __launch_bounds__(256) col2im_nhwc_fast_kernel(const bf16* __restrict__ dcol, const TA* __restrict__ add, long ldadd,
                                                               TO* __restrict__ dx, long lddx, int C, int k, int s, int p, int Ho, int Wo,
                                                               unsigned total, Col2imDivs dv) {
  pdl_trigger();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  unsigned pix, cg, q, ix, b, iy;
  fdivmod(idx, dv.c8, pix, cg);
  fdivmod(pix, dv.W, q, ix);
  fdivmod(q, dv.H, b, iy);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  if (add) load8(add + (long)pix * ldadd + cg * 8, acc);
  // only taps with (iy + p - kh) divisible by the stride contribute (one tap for the non-overlapping SR convolutions k = s = R)
  unsigned qy, ry, qx, rx;
  fdivmod(iy + (unsigned)p, dv.s, qy, ry);   // iy + p = qy * s + ry: tap kh = ry + j*s reads output row qy - j
  fdivmod(ix + (unsigned)p, dv.s, qx, rx);
  for (int kh = (int)ry, oy = (int)qy; kh < k && oy >= 0; kh += s, oy--) {
    if (oy >= Ho) continue;
    for (int kw = (int)rx, ox = (int)qx; kw < k && ox >= 0; kw += s, ox--) {
      if (ox >= Wo) continue;
      float v[8];
      load8(dcol + ((((long)b * Ho + oy) * Wo + ox) * (k * k) + kh * k + kw) * C + cg * 8, v);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[i] += v[i];
    }
  }
  store8(dx + (long)pix * lddx + cg * 8, acc);
}
CMX_API int cmx_col2im_nhwc(const void* dcol, const void* add, int add_dtype, int64_t ldadd, void* dx, int dx_dtype, int64_t lddx,
                            int B, int H, int W, int C, int k, int s, int p, int Ho, int Wo, void* stream) {
  CMX_REQUIRE(C % 8 == 0 && lddx % 8 == 0, "col2im_nhwc: C %% 8");
  const long total = (long)B * H * W * (C >> 3);
  if (total == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(total, 256));
#define C2I(TA, TO)                                                                                                            \
  do {                                                                                                                         \
    if (total < (1L << 31)) {                                                                                                  \
      Col2imDivs dv{make_fastdiv((unsigned)(C >> 3)), make_fastdiv((unsigned)W), make_fastdiv((unsigned)H), make_fastdiv((unsigned)s)}; \
      col2im_nhwc_fast_kernel<TA, TO><<<grid, 256, 0, st>>>((const bf16*)dcol, (const TA*)add, ldadd, (TO*)dx, lddx, C, k, s, p, Ho, \
                                                            Wo, (unsigned)total, dv);                                             \
    }                                                                                                                          \
    else                                                                                                                       \
      col2im_nhwc_kernel<TA, TO, long><<<grid, 256, 0, st>>>((const bf16*)dcol, (const TA*)add, ldadd, (TO*)dx, lddx, B, H, W,  \
                                                             C, k, s, p, Ho, Wo);                                               \
  } while (0)
  if (add_dtype == CMX_F32 && dx_dtype == CMX_F32) C2I(float, float);
  else if (add_dtype == CMX_F32) C2I(float, bf16);
  else if (dx_dtype == CMX_F32) C2I(bf16, float);
  else C2I(bf16, bf16);
#undef C2I
  LAUNCH_DONE("col2im_nhwc");
}

// ---- conv weight pack / grad unpack ------------------------------------------------------------------
__global__ void convw_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int Co, int Ci, int kh, int kw, int kpad) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Co * kpad) return;
  const int j = (int)(idx % kpad);
  const int co = (int)(idx / kpad);
  float v = 0.f;
  if (j < kh * kw * Ci) {
    const int ci = j % Ci, tap = j / Ci;
    v = w[((long)co * Ci + ci) * (kh * kw) + tap];
  }
  wp[idx] = __float2bfloat16(v);
}
CMX_API int cmx_convw_pack(const float* w, void* wp, int Co, int Ci, int kh, int kw, int kpad, void* stream) {
  const long n = (long)Co * kpad;
  convw_pack_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)wp, Co, Ci, kh, kw, kpad);
  LAUNCH_DONE("convw_pack");
}
__global__ void convw_unpack_grad_kernel(const float* __restrict__ gp, float* __restrict__ gw, int Co, int Ci, int kh, int kw, int kpad) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = kh * kw * Ci;
  if (idx >= (long)Co * K) return;
  const int j = (int)(idx % K);
  const int co = (int)(idx / K);
  const int ci = j % Ci, tap = j / Ci;
  gw[((long)co * Ci + ci) * (kh * kw) + tap] += gp[(long)co * kpad + j];
}
CMX_API int cmx_convw_unpack_grad(const float* gp, float* gw, int Co, int Ci, int kh, int kw, int kpad, void* stream) {
  const long n = (long)Co * kh * kw * Ci;
  convw_unpack_grad_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(gp, gw, Co, Ci, kh, kw, kpad);
  LAUNCH_DONE("convw_unpack_grad");
}

// all real convolutions of the model in ONE launch each way (34 for MiT-B2): blockIdx.y = convolution, blockIdx.x walks
// the output channels; one [Ci, kh*kw] <-> [(kh*kw), Ci] row transpose per CTA through shared memory, so both the
// global read and the global write of a row are contiguous
constexpr int CONVW_MAXK = 4608;  // 512 x 3 x 3 (the largest patch embed); SR convs: C * R * R <= 4096
__global__ void __launch_bounds__(256) convw_pack_multi_kernel(const CmxConvDesc* __restrict__ d) {
  pdl_trigger();
  __shared__ float row[CONVW_MAXK];
  const CmxConvDesc c = d[blockIdx.y];
  const int taps = c.kh * c.kw, K = taps * c.Ci;
  bf16* wp = reinterpret_cast<bf16*>(c.wp);
  for (int co = blockIdx.x; co < c.Co; co += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) row[i] = c.w[(long)co * K + i];   // [ci][tap], contiguous
    __syncthreads();
    for (int j = threadIdx.x; j < c.kpad; j += blockDim.x) {
      float v = 0.f;
      if (j < K) v = row[(j % c.Ci) * taps + j / c.Ci];
      wp[(long)co * c.kpad + j] = __float2bfloat16(v);
    }
  }
}
__global__ void __launch_bounds__(256) convw_unpack_multi_kernel(const CmxConvDesc* __restrict__ d) {
  pdl_trigger();
  __shared__ float row[CONVW_MAXK];
  const CmxConvDesc c = d[blockIdx.y];
  const int taps = c.kh * c.kw, K = taps * c.Ci;
  for (int co = blockIdx.x; co < c.Co; co += gridDim.x) {
    __syncthreads();
    for (int j = threadIdx.x; j < K; j += blockDim.x) row[j] = c.gp[(long)co * c.kpad + j];   // [tap][ci], contiguous
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) c.gw[(long)co * K + i] += row[(i % taps) * c.Ci + i / taps];
  }
}
static int convw_check(const CmxConvDesc*, int n) {
  CMX_REQUIRE(n <= 65535, "convw_*_multi: too many convolutions");
  return 0;
}
CMX_API int cmx_convw_pack_multi(const CmxConvDesc* descs, int n, void* stream) {
  if (n <= 0) return 0;
  if (int rc = convw_check(descs, n)) return rc;
  convw_pack_multi_kernel<<<dim3(128, n), 256, 0, (cudaStream_t)stream>>>(descs);
  LAUNCH_DONE("convw_pack_multi");
}
CMX_API int cmx_convw_unpack_grad_multi(const CmxConvDesc* descs, int n, void* stream) {
  if (n <= 0) return 0;
  if (int rc = convw_check(descs, n)) return rc;
  convw_unpack_multi_kernel<<<dim3(128, n), 256, 0, (cudaStream_t)stream>>>(descs);
  LAUNCH_DONE("convw_unpack_grad_multi");
}

// ---- casts / axpby -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long n) {
  pdl_trigger();
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float f[8];
    load8(x + i, f);
    store8(y + i, f);
  } else {
    for (long j = i; j < n; j++) y[j] = __float2bfloat16(x[j]);
  }
}
CMX_API int cmx_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  CMX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "cast: pointers must be 16B aligned");
  cast_f32_bf16_kernel<<<cdiv(cdiv(n, 8), 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, n);
  LAUNCH_DONE("cast_f32_bf16");
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, long n) {
  pdl_trigger();
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float f[8];
    load8(x + i, f);
    store8(y + i, f);
  } else {
    for (long j = i; j < n; j++) y[j] = __bfloat162float(x[j]);
  }
}
CMX_API int cmx_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  CMX_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "cast: pointers must be 16B aligned");
  cast_bf16_f32_kernel<<<cdiv(cdiv(n, 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, y, n);
  LAUNCH_DONE("cast_bf16_f32");
}
__global__ void __launch_bounds__(256) axpby_kernel(float a, const float* __restrict__ x, float b, const float* __restrict__ y,
                                                    float* __restrict__ out, long n) {
  pdl_trigger();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = a * x[i];
  if (y) v += b * y[i];
  out[i] = v;
}
CMX_API int cmx_axpby_f32(float a, const float* x, float b, const float* y, float* out, int64_t n, void* stream) {
  if (n == 0) return 0;
  axpby_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a, x, b, y, out, n);
  LAUNCH_DONE("axpby");
}

// ---- column sum (bias gradients): out[n] += sum_m x[m,n] ----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long ldx, float* __restrict__ out, long M, int N,
                                                     int rows_per_cta, long ogs) {
  pdl_trigger();
  x += (long)blockIdx.z * M * ldx;   // grouped launch: stacked row blocks, outputs ogs elements apart
  out += (long)blockIdx.z * ogs;
  __shared__ float s1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a = 0.f;
  if (c < N)
    for (long r = r0 + ty; r < r1; r += 8) a += ld1(x + r * ldx + c);
  s1[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += s1[i][tx];
    atomicAdd(out + c, t);
  }
}
// 8 columns (one 16-byte load) per thread; block = NG column groups x (256/NG) row lanes
__global__ void __launch_bounds__(256) colsum_vec8_kernel(const bf16* __restrict__ x, long ldx, float* __restrict__ out, long M, int N,
                                                          int ng, int rows_per_cta, long ogs) {
  pdl_trigger();
  x += (long)blockIdx.z * M * ldx;   // grouped launch: stacked row blocks, outputs ogs elements apart
  out += (long)blockIdx.z * ogs;
  __shared__ float sacc[256][9];
  const int tid = threadIdx.x;
  const int cg = tid % ng, rl = tid / ng, nrl = 256 / ng;
  const int c = (blockIdx.x * ng + cg) * 8;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 0.f;
  if (c < N && rl < nrl) {
    long r = r0 + rl;
    for (; r + 3 * nrl < r1; r += 4 * nrl) {  // 4 independent 16-byte loads in flight
      float v0[8], v1[8], v2[8], v3[8];
      load8(x + r * ldx + c, v0);
      load8(x + (r + nrl) * ldx + c, v1);
      load8(x + (r + 2 * nrl) * ldx + c, v2);
      load8(x + (r + 3 * nrl) * ldx + c, v3);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] += (v0[i] + v1[i]) + (v2[i] + v3[i]);
    }
    for (; r < r1; r += nrl) {
      float v[8];
      load8(x + r * ldx + c, v);
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) sacc[tid][i] = a[i];
  __syncthreads();
  if (tid < ng * 8) {
    const int g = tid >> 3, i = tid & 7;
    const int cc = (blockIdx.x * ng + g) * 8 + i;
    if (cc < N) {
      float t = 0.f;
      for (int l = 0; l < nrl; l++) t += sacc[l * ng + g][i];
      atomicAdd(out + cc, t);
    }
  }
}

CMX_API int cmx_colsum(const void* x, int x_dtype, int64_t ldx, float* out, int64_t M, int N, int groups, int64_t out_gs,
                       void* stream) {
  if (M == 0 || N == 0) return 0;
  CMX_REQUIRE(groups >= 1 && groups <= 65535, "colsum: groups=%d", groups);
  const long ogs = out_gs;
  if (x_dtype == CMX_BF16 && N % 8 == 0 && ldx % 8 == 0 && (((uintptr_t)x) & 15) == 0) {
    int ng = N / 8;
    if (ng > 32) ng = 32;
    while (256 % ng) ng--;  // ng must divide 256 (N/8 in {4,8,16,20->16,32,...})
    const int nrl = 256 / ng;
    long want_ctas = 148L * 4 / groups;
    const int gx = cdiv(N / 8, ng);
    long rows_per_cta = (M * gx + want_ctas - 1) / want_ctas;
    rows_per_cta = (rows_per_cta + nrl - 1) / nrl * nrl;
    if (rows_per_cta < 4L * nrl) rows_per_cta = 4L * nrl;
    dim3 grid(gx, cdiv(M, rows_per_cta), groups);
    colsum_vec8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, out, M, N, ng, (int)rows_per_cta, ogs);
    LAUNCH_DONE("colsum_vec8");
  }
  const int rows_per_cta = 512;
  dim3 grid(cdiv(N, 32), cdiv(M, rows_per_cta), groups);
  CMX_REQUIRE(grid.y <= 65535, "colsum: M too large");
  if (x_dtype == CMX_F32) colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, out, M, N, rows_per_cta, ogs);
  else colsum_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, out, M, N, rows_per_cta, ogs);
  LAUNCH_DONE("colsum");
}

// ---- ReLU backward in place ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_bwd_kernel(bf16* __restrict__ dy, long lddy, const bf16* __restrict__ y, long ldy, long M,
                                                       int N) {
  pdl_trigger();
  const int n8 = N >> 3;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * n8) return;
  const long row = idx / n8;
  const int c = (int)(idx % n8) * 8;
  float d[8], v[8];
  load8(dy + row * lddy + c, d);
  load8(y + row * ldy + c, v);
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = v[i] > 0.f ? d[i] : 0.f;
  store8(dy + row * lddy + c, d);
}
CMX_API int cmx_relu_bwd(void* dy, int64_t lddy, const void* y, int64_t ldy, int64_t M, int N, void* stream) {
  CMX_REQUIRE(N % 8 == 0 && lddy % 8 == 0 && ldy % 8 == 0, "relu_bwd: N %% 8");
  if (M == 0) return 0;
  relu_bwd_kernel<<<cdiv(M * (N >> 3), 256), 256, 0, (cudaStream_t)stream>>>((bf16*)dy, lddy, (const bf16*)y, ldy, M, N);
  LAUNCH_DONE("relu_bwd");
}

// ---- row softmax (self-attention probabilities); one warp per row, n <= 1024 ----------------------------
constexpr int SM_MAXJ = 32;
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(const float* __restrict__ s, long lds, bf16* __restrict__ p, long ldp,
                                                               long rows, int n) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * lds;
  float v[SM_MAXJ];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < SM_MAXJ; j++) {
    const int c = lane + 32 * j;
    v[j] = c < n ? sr[c] : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < SM_MAXJ; j++) {
    const int c = lane + 32 * j;
    v[j] = c < n ? __expf(v[j] - mx) : 0.f;
    sum += v[j];
  }
  const float inv = 1.f / warp_sum(sum);
  bf16* pr = p + row * ldp;
#pragma unroll
  for (int j = 0; j < SM_MAXJ; j++) {
    const int c = lane + 32 * j;
    if (c < n) pr[c] = __float2bfloat16(v[j] * inv);
  }
}
CMX_API int cmx_softmax_rows_fwd(const float* s, int64_t lds, void* p, int64_t ldp, int64_t rows, int n, void* stream) {
  CMX_REQUIRE(n > 0 && n <= 32 * SM_MAXJ, "softmax_rows: n=%d unsupported", n);
  if (rows == 0) return 0;
  softmax_rows_fwd_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(s, lds, (bf16*)p, ldp, rows, n);
  LAUNCH_DONE("softmax_rows_fwd");
}
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const bf16* __restrict__ p, long ldp, const float* __restrict__ dp,
                                                               long lddp, float scale, bf16* __restrict__ ds, long ldds, long rows,
                                                               int n) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float pv[SM_MAXJ], dv[SM_MAXJ];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < SM_MAXJ; j++) {
    const int c = lane + 32 * j;
    pv[j] = c < n ? __bfloat162float(p[row * ldp + c]) : 0.f;
    dv[j] = c < n ? dp[row * lddp + c] : 0.f;
    dot += pv[j] * dv[j];
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < SM_MAXJ; j++) {
    const int c = lane + 32 * j;
    if (c < n) ds[row * ldds + c] = __float2bfloat16(scale * pv[j] * (dv[j] - dot));
  }
}
CMX_API int cmx_softmax_rows_bwd(const void* p, int64_t ldp, const float* dp, int64_t lddp, float scale, void* ds, int64_t ldds,
                                 int64_t rows, int n, void* stream) {
  CMX_REQUIRE(n > 0 && n <= 32 * SM_MAXJ, "softmax_rows_bwd: n=%d unsupported", n);
  if (rows == 0) return 0;
  softmax_rows_bwd_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)p, ldp, dp, lddp, scale, (bf16*)ds, ldds, rows, n);
  LAUNCH_DONE("softmax_rows_bwd");
}

// ---- softmax over dim -2 of [nb, d, d] (FFM context, d = 64): thread per column ------------------------
__global__ void softmax_dim2_fwd_kernel(const float* __restrict__ c, float scale, float* __restrict__ p32, bf16* __restrict__ p16, int d) {
  pdl_trigger();
  const int col = threadIdx.x;
  const long base = (long)blockIdx.x * d * d;
  if (col >= d) return;
  float mx = -INFINITY;
  for (int i = 0; i < d; i++) mx = fmaxf(mx, c[base + (long)i * d + col] * scale);
  float sum = 0.f;
  for (int i = 0; i < d; i++) sum += __expf(c[base + (long)i * d + col] * scale - mx);
  const float inv = 1.f / sum;
  for (int i = 0; i < d; i++) {
    const float v = __expf(c[base + (long)i * d + col] * scale - mx) * inv;
    p32[base + (long)i * d + col] = v;
    p16[base + (long)i * d + col] = __float2bfloat16(v);
  }
}
CMX_API int cmx_softmax_dim2_fwd(const float* c, float scale, float* p32, void* p16, int nb, int d, void* stream) {
  CMX_REQUIRE(d > 0 && d <= 1024, "softmax_dim2: d");
  if (nb == 0) return 0;
  softmax_dim2_fwd_kernel<<<nb, ((d + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(c, scale, p32, (bf16*)p16, d);
  LAUNCH_DONE("softmax_dim2_fwd");
}
__global__ void softmax_dim2_bwd_kernel(const float* __restrict__ p32, const float* __restrict__ dp, float scale, bf16* __restrict__ dc16, int d) {
  pdl_trigger();
  const int col = threadIdx.x;
  const long base = (long)blockIdx.x * d * d;
  if (col >= d) return;
  float dot = 0.f;
  for (int i = 0; i < d; i++) dot += p32[base + (long)i * d + col] * dp[base + (long)i * d + col];
  for (int i = 0; i < d; i++) {
    const long o = base + (long)i * d + col;
    dc16[o] = __float2bfloat16(scale * p32[o] * (dp[o] - dot));
  }
}
CMX_API int cmx_softmax_dim2_bwd(const float* p32, const float* dp, float scale, void* dc16, int nb, int d, void* stream) {
  CMX_REQUIRE(d > 0 && d <= 1024, "softmax_dim2_bwd: d");
  if (nb == 0) return 0;
  softmax_dim2_bwd_kernel<<<nb, ((d + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(p32, dp, scale, (bf16*)dc16, d);
  LAUNCH_DONE("softmax_dim2_bwd");
}
