// LayerNorm (fwd/bwd) and BatchNorm (batch-stat reduction, apply, bwd) on token-major [M,C] matrices.
// Memory-bound kernels: one warp per row for LayerNorm (row cached in registers, 8/16-byte accesses),
// channel-coalesced column reductions with double atomics for BatchNorm statistics.
#include "common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>
#include <initializer_list>
extern std::atomic<long long> g_cmx_launches;

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: C % 4 == 0, C <= 1024.  Lane l owns elements {4*(l + 32*j) .. +3}.
// ------------------------------------------------------------------------------------------------
constexpr int LN_MAXJ = 4;  // 4 * 32 lanes * 4 = 512 channels (max embed dim of MiT-B0..B5)

template <typename TX, typename TY>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, long ldx, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, TY* __restrict__ y, long ldy,
                                                     float* __restrict__ mean, float* __restrict__ rstd, long M, int C) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const TX* xr = x + row * ldx;
  float v[LN_MAXJ][4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
      load4(xr + c, v[j]);
      s += v[j][0] + v[j][1] + v[j][2] + v[j][3];
    }
  }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float d = v[j][i] - mu;
        q += d * d;
      }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  TY* yr = y + row * ldy;
#pragma unroll
  for (int j = 0; j < LN_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
      float g[4], b[4], o[4];
      load4(gamma + c, g);
      load4(beta + c, b);
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = (v[j][i] - mu) * rs * g[i] + b[i];
      store4(yr + c, o);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// LayerNorm v2 (production path for C % 8 == 0): G lanes cooperate on one row, 32/G rows per warp, every lane
// owns J vectors of 8 consecutive channels (16-byte bf16 / 32-byte fp32 accesses).  C = 64 -> 4 rows per warp,
// C = 128 -> 2, C >= 160 -> 1; reductions are xor-shuffles inside the G-lane group.
// ------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TX, typename TY, int G, int J>
__global__ void __launch_bounds__(256) ln_fwd_v2_kernel(const TX* __restrict__ x, long ldx, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, TY* __restrict__ y, long ldy,
                                                        float* __restrict__ mean, float* __restrict__ rstd, long M, int C, long pgs,
                                                        float inv_c) {
  pdl_trigger();
  if (gridDim.y > 1) {  // grouped launch: group g = rows [g*M, (g+1)*M) of the stacked tensors, parameters pgs elements apart
    const long g = blockIdx.y;
    x += g * M * ldx; y += g * M * ldy; gamma += g * pgs; beta += g * pgs;
    if (mean) mean += g * M;
    if (rstd) rstd += g * M;
  }
  // Persistent row walk: the per-thread set-up (index arithmetic, gamma / beta) is paid once per thread, not once per row - the
  // one-row-per-thread form spent ~230 instructions on 8 elements and was issue bound (ncu: 70 % issue activity at 3.3 TB/s)
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int lg = lane % G;
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float gm[J][8], bt[J][8];
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int c = 8 * (lg + G * j);
#pragma unroll
    for (int i = 0; i < 8; i++) { gm[j][i] = 0.f; bt[j][i] = 0.f; }
    if (c < C) { load8(gamma + c, gm[j]); load8(beta + c, bt[j]); }
  }
  const long nslots = (M + RPW - 1) / RPW;
  for (long slot = (long)blockIdx.x * nwarp + warp; slot < nslots; slot += (long)gridDim.x * nwarp) {
    const long row = slot * RPW + lane / G;
    const bool ok = row < M;
    float v[J][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
#pragma unroll
      for (int i = 0; i < 8; i++) v[j][i] = 0.f;
      if (ok && c < C) {
        load8(x + row * ldx + c, v[j]);
#pragma unroll
        for (int i = 0; i < 8; i++) s += v[j][i];
      }
    }
    const float mu = group_sum<G>(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float d = v[j][i] - mu;
          q += d * d;
        }
      }
    }
    const float rs = rsqrtf(group_sum<G>(q) * inv_c + eps);
    if (!ok) continue;
    if (lg == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
      if (c < C) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = (v[j][i] - mu) * rs * gm[j][i] + bt[j][i];
        store8(y + row * ldy + c, o);
      }
    }
  }
}

template <typename TDY, typename TX, typename TDX, int G, int J>
__global__ void __launch_bounds__(256, (J == 1 ? 3 : 2)) ln_bwd_v2_kernel(const TDY* __restrict__ dy, long lddy, const bf16* __restrict__ dy2, long lddy2,
                                                        const TX* __restrict__ x, long ldx, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                        const float* __restrict__ dres, long lddres, TDX* __restrict__ dx, long lddx,
                                                        bf16* __restrict__ dxbf, long lddxbf, const float* __restrict__ scale,
                                                        int rows_per_sample, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                        float* __restrict__ dbias, long M, int C, long pgs, long sgs) {
  pdl_trigger();
  if (gridDim.y > 1) {  // grouped launch (see ln_fwd_v2_kernel); scale: sgs elements apart
    const long g = blockIdx.y;
    dy += g * M * lddy; x += g * M * ldx; mean += g * M; rstd += g * M; gamma += g * pgs;
    if (dy2) dy2 += g * M * lddy2;
    if (dres) dres += g * M * lddres;
    if (dx) dx += g * M * lddx;
    if (dxbf) dxbf += g * M * lddxbf;
    if (scale) scale += g * sgs;
    if (dgamma) { dgamma += g * pgs; dbeta += g * pgs; }
    if (dbias) dbias += g * pgs;
  }
  constexpr int RPW = 32 / G;
  __shared__ __align__(16) float sh_red[8][512];   // per-warp column partial sums (blockDim.x == 256)
  const int lane = threadIdx.x & 31;
  const int lg = lane % G;
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float ag[J][8], ab[J][8], gm[J][8], as[J][8];
#pragma unroll
  for (int j = 0; j < J; j++) {
    const int c = 8 * (lg + G * j);
#pragma unroll
    for (int i = 0; i < 8; i++) { ag[j][i] = 0.f; ab[j][i] = 0.f; gm[j][i] = 0.f; as[j][i] = 0.f; }
    if (c < C) load8(gamma + c, gm[j]);
  }
  const long nslots = (M + RPW - 1) / RPW;  // warp-iterations
  const float inv_c = 1.f / (float)C;
  for (long slot = (long)blockIdx.x * nwarp + warp; slot < nslots; slot += (long)gridDim.x * nwarp) {
    const long row = slot * RPW + lane / G;
    const bool ok = row < M;
    const float mu = ok ? mean[row] : 0.f, rs = ok ? rstd[row] : 0.f;
    float g[J][8], xh[J][8], rr[J][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
#pragma unroll
      for (int i = 0; i < 8; i++) { g[j][i] = 0.f; xh[j][i] = 0.f; rr[j][i] = 0.f; }
      if (ok && c < C) {
        float d[8], xv[8];
        load8(dy + row * lddy + c, d);
        if (dres) load8(dres + row * lddres + c, rr[j]);  // issued with the other loads: one memory round trip per row
        if (dy2) {
          float d2[8];
          load8(dy2 + row * lddy2 + c, d2);
#pragma unroll
          for (int i = 0; i < 8; i++) d[i] += d2[i];
        }
        load8(x + row * ldx + c, xv);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          xh[j][i] = (xv[i] - mu) * rs;
          ag[j][i] += d[i] * xh[j][i];
          ab[j][i] += d[i];
          g[j][i] = d[i] * gm[j][i];
          s1 += g[j][i];
          s2 += g[j][i] * xh[j][i];
        }
      }
    }
    s1 = group_sum<G>(s1) * inv_c;   // (a per-row IEEE division is ~10 instructions; the kernel walks millions of rows)
    s2 = group_sum<G>(s2) * inv_c;
    if (!ok) continue;
    const float sc = (scale && dxbf) ? scale[(int)(row / rows_per_sample)] : 1.f;
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
      if (c < C) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = rs * (g[j][i] - s1 - xh[j][i] * s2);
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] += rr[j][i];
        if (dx) store8(dx + row * lddx + c, o);
        if (dxbf) {
#pragma unroll
          for (int i = 0; i < 8; i++) o[i] *= sc;
          store8(dxbf + row * lddxbf + c, o);
        }
        if (dbias) {
#pragma unroll
          for (int i = 0; i < 8; i++) as[j][i] += o[i];
        }
      }
    }
  }
  // column reductions: fold the 32/G row groups of each warp with shuffles, park the per-warp sums in a shared slab
  // (plain 16-byte stores, no shared-memory atomics), add the 8 warps per column, one global atomic per column per CTA
  auto reduce_cols = [&](float (&acc)[J][8], float* __restrict__ gout) {
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 8 * (lg + G * j);
      float r[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float a = acc[j][i];
#pragma unroll
        for (int o = G; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        r[i] = a;
      }
      if (lane < G && c < C) store8(&sh_red[warp][c], r);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) t += sh_red[w][i];
      atomicAdd(gout + i, t);
    }
    __syncthreads();
  };
  if (dbias) reduce_cols(as, dbias);
  if (dgamma) {
    reduce_cols(ag, dgamma);
    reduce_cols(ab, dbeta);
  }
}

static inline void ln_gj(int C, int& G, int& J) {
  const int nvec = C / 8;
  G = nvec <= 4 ? 4 : nvec <= 8 ? 8 : nvec <= 16 ? 16 : 32;
  J = (nvec + G - 1) / G;
}
#define LN_GJ_DISPATCH(MACRO)                     \
  do {                                            \
    if (G == 4) MACRO(4, 1);                      \
    else if (G == 8) MACRO(8, 1);                 \
    else if (G == 16) MACRO(16, 1);               \
    else if (J == 1) MACRO(32, 1);                \
    else MACRO(32, 2);                            \
  } while (0)

static size_t esz(int dtype) { return dtype == CMX_F32 ? 4 : 2; }

CMX_API int cmx_layernorm_fwd(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, float eps,
                              void* y, int y_dtype, int64_t ldy, float* mean, float* rstd, int64_t M, int C, int groups,
                              int64_t param_gs, void* stream) {
  CMX_REQUIRE(C % 4 == 0 && C <= 4 * 32 * LN_MAXJ && C > 0, "layernorm: C=%d unsupported", C);
  CMX_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "layernorm: ld must be a multiple of 4");
  CMX_REQUIRE(groups >= 1 && groups <= 65535, "layernorm: groups=%d", groups);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = C % 8 == 0 && C <= 512 && ldx % 8 == 0 && ldy % 8 == 0 && (groups == 1 || param_gs % 4 == 0);
  if (groups > 1 && !vec) {   // the scalar kernels are single-group: one launch per group
    for (int g = 0; g < groups; g++) {
      int rc = cmx_layernorm_fwd((const char*)x + (size_t)g * M * ldx * esz(x_dtype), x_dtype, ldx, gamma + g * param_gs, beta + g * param_gs,
                                 eps, (char*)y + (size_t)g * M * ldy * esz(y_dtype), y_dtype, ldy, mean ? mean + g * M : nullptr,
                                 rstd ? rstd + g * M : nullptr, M, C, 1, 0, stream);
      if (rc) return rc;
    }
    return 0;
  }
  if (vec) {
    int G, J;
    ln_gj(C, G, J);
    const int rpb = 8 * (32 / G);
    long gx = cdiv(M, rpb);
    const long resident = 148L * 6 / groups;   // ~6 CTAs of 256 threads per SM; the rows beyond are walked grid-stride
    if (gx > resident) gx = resident;
    if (gx < 1) gx = 1;
    dim3 grid2((unsigned)gx, groups);
    const long pgs = param_gs;
    const float inv_c = 1.f / (float)C;
#define LN_F2T(TX, TY, Gv, Jv) ln_fwd_v2_kernel<TX, TY, Gv, Jv><<<grid2, 256, 0, st>>>((const TX*)x, ldx, gamma, beta, eps, (TY*)y, ldy, mean, rstd, M, C, pgs, inv_c)
#define LN_F2(Gv, Jv)                                                                 \
  do {                                                                                \
    if (x_dtype == CMX_F32 && y_dtype == CMX_BF16) LN_F2T(float, bf16, Gv, Jv);       \
    else if (x_dtype == CMX_F32 && y_dtype == CMX_F32) LN_F2T(float, float, Gv, Jv);  \
    else if (x_dtype == CMX_BF16 && y_dtype == CMX_BF16) LN_F2T(bf16, bf16, Gv, Jv);  \
    else LN_F2T(bf16, float, Gv, Jv);                                                 \
  } while (0)
    LN_GJ_DISPATCH(LN_F2);
#undef LN_F2
#undef LN_F2T
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("ln_fwd_v2");
    return 0;
  }
  const int wpb = 8;
  dim3 grid(cdiv(M, wpb));
#define LN_F(TX, TY) ln_fwd_kernel<TX, TY><<<grid, wpb * 32, 0, st>>>((const TX*)x, ldx, gamma, beta, eps, (TY*)y, ldy, mean, rstd, M, C)
  if (x_dtype == CMX_F32 && y_dtype == CMX_BF16) LN_F(float, bf16);
  else if (x_dtype == CMX_F32 && y_dtype == CMX_F32) LN_F(float, float);
  else if (x_dtype == CMX_BF16 && y_dtype == CMX_BF16) LN_F(bf16, bf16);
  else LN_F(bf16, float);
#undef LN_F
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("ln_fwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  Each CTA walks rows grid-stride, each warp one row at a time; per-lane
// dgamma/dbeta partials stay in registers, are combined across the CTA's warps in shared memory and
// flushed with one atomicAdd per channel per CTA.
// ------------------------------------------------------------------------------------------------
template <typename TDY, typename TX, typename TDX>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TDY* __restrict__ dy, long lddy, const bf16* __restrict__ dy2, long lddy2,
                                                     const TX* __restrict__ x, long ldx, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                     const float* __restrict__ dres, long lddres, TDX* __restrict__ dx, long lddx,
                                                     bf16* __restrict__ dxbf, long lddxbf, const float* __restrict__ scale,
                                                     int rows_per_sample, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     long M, int C) {
  pdl_trigger();
  __shared__ float sh_g[1024];
  __shared__ float sh_b[1024];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sh_g[i] = 0.f; sh_b[i] = 0.f; }
  __syncthreads();
  float ag[LN_MAXJ][4], ab[LN_MAXJ][4], gm[LN_MAXJ][4];
#pragma unroll
  for (int j = 0; j < LN_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
#pragma unroll
    for (int i = 0; i < 4; i++) { ag[j][i] = 0.f; ab[j][i] = 0.f; gm[j][i] = 0.f; }
    if (c < C) load4(gamma + c, gm[j]);
  }
  for (long row = (long)blockIdx.x * nwarp + warp; row < M; row += (long)gridDim.x * nwarp) {
    const float mu = mean[row], rs = rstd[row];
    float g[LN_MAXJ][4], xh[LN_MAXJ][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; j++) {
      const int c = 4 * (lane + 32 * j);
      if (c < C) {
        float d[4], xv[4];
        load4(dy + row * lddy + c, d);
        if (dy2) {
          float d2[4];
          load4(dy2 + row * lddy2 + c, d2);
#pragma unroll
          for (int i = 0; i < 4; i++) d[i] += d2[i];
        }
        load4(x + row * ldx + c, xv);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          xh[j][i] = (xv[i] - mu) * rs;
          ag[j][i] += d[i] * xh[j][i];
          ab[j][i] += d[i];
          g[j][i] = d[i] * gm[j][i];
          s1 += g[j][i];
          s2 += g[j][i] * xh[j][i];
        }
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    const float sc = (scale && dxbf) ? scale[row / rows_per_sample] : 1.f;
#pragma unroll
    for (int j = 0; j < LN_MAXJ; j++) {
      const int c = 4 * (lane + 32 * j);
      if (c < C) {
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = rs * (g[j][i] - s1 - xh[j][i] * s2);
        if (dres) {
          float r[4];
          load4(dres + row * lddres + c, r);
#pragma unroll
          for (int i = 0; i < 4; i++) o[i] += r[i];
        }
        if (dx) store4(dx + row * lddx + c, o);
        if (dxbf) {
#pragma unroll
          for (int i = 0; i < 4; i++) o[i] *= sc;
          store4(dxbf + row * lddxbf + c, o);
        }
      }
    }
  }
  if (dgamma) {
#pragma unroll
    for (int j = 0; j < LN_MAXJ; j++) {
      const int c = 4 * (lane + 32 * j);
      if (c < C) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          atomicAdd(&sh_g[c + i], ag[j][i]);
          atomicAdd(&sh_b[c + i], ab[j][i]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(dgamma + i, sh_g[i]);
      atomicAdd(dbeta + i, sh_b[i]);
    }
  }
}

CMX_API int cmx_layernorm_bwd(const void* dy, int dy_dtype, int64_t lddy, const void* dy2, int64_t lddy2, const void* x,
                              int x_dtype, int64_t ldx, const float* mean, const float* rstd, const float* gamma,
                              const float* dres, int64_t lddres, void* dx, int dx_dtype, int64_t lddx, void* dx_bf,
                              int64_t lddxbf, const float* scale, int rows_per_sample, float* dgamma, float* dbeta,
                              float* dbias, int64_t M, int C, int groups, int64_t param_gs, int64_t scale_gs, void* stream) {
  CMX_REQUIRE(C % 4 == 0 && C <= 4 * 32 * LN_MAXJ && C > 0, "layernorm_bwd: C=%d unsupported", C);
  CMX_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma/dbeta must come together");
  CMX_REQUIRE(groups >= 1 && groups <= 65535, "layernorm_bwd: groups=%d", groups);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_per_sample <= 0) rows_per_sample = 1;
  const bool vec = C % 8 == 0 && C <= 512 && lddy % 8 == 0 && ldx % 8 == 0 && (!dy2 || lddy2 % 8 == 0) && (!dres || lddres % 8 == 0) &&
                   (!dx || lddx % 8 == 0) && (!dx_bf || lddxbf % 8 == 0) && (groups == 1 || param_gs % 4 == 0);
  if (groups > 1 && !vec) {   // the scalar kernel is single-group: one launch per group
    for (int g = 0; g < groups; g++) {
      const size_t r = (size_t)g * M;
      int rc = cmx_layernorm_bwd((const char*)dy + r * lddy * esz(dy_dtype), dy_dtype, lddy, dy2 ? (const char*)dy2 + r * lddy2 * 2 : nullptr,
                                 lddy2, (const char*)x + r * ldx * esz(x_dtype), x_dtype, ldx, mean + r, rstd + r, gamma + g * param_gs,
                                 dres ? dres + r * lddres : nullptr, lddres, dx ? (char*)dx + r * lddx * esz(dx_dtype) : nullptr, dx_dtype,
                                 lddx, dx_bf ? (char*)dx_bf + r * lddxbf * 2 : nullptr, lddxbf, scale ? scale + g * scale_gs : nullptr,
                                 rows_per_sample, dgamma ? dgamma + g * param_gs : nullptr, dbeta ? dbeta + g * param_gs : nullptr,
                                 dbias ? dbias + g * param_gs : nullptr, M, C, 1, 0, 0, stream);
      if (rc) return rc;
    }
    return 0;
  }
  if (vec) {
    int G, J;
    ln_gj(C, G, J);
    int gx = cdiv(M, 8 * (32 / G));
    const int resident = 148 * (J == 1 ? 3 : 2) / groups;  // one full wave of CTAs (see __launch_bounds__), grid-stride beyond
    if (gx > resident) gx = resident;
    if (gx < 1) gx = 1;
    dim3 grid2(gx, groups);
    const long pgs = param_gs, sgs = scale_gs;
#define LN_B2T(TDY, TX, TDX, Gv, Jv)                                                                                            \
  ln_bwd_v2_kernel<TDY, TX, TDX, Gv, Jv><<<grid2, 256, 0, st>>>((const TDY*)dy, lddy, (const bf16*)dy2, lddy2, (const TX*)x, ldx, \
                                                                mean, rstd, gamma, dres, lddres, (TDX*)dx, lddx, (bf16*)dx_bf,    \
                                                                lddxbf, scale, rows_per_sample, dgamma, dbeta, dbias, M, C, pgs, sgs)
#define LN_B2(Gv, Jv)                                                            \
  do {                                                                           \
    switch (dy_dtype * 4 + x_dtype * 2 + dx_dtype) {                             \
      case 0: LN_B2T(bf16, bf16, bf16, Gv, Jv); break;                           \
      case 1: LN_B2T(bf16, bf16, float, Gv, Jv); break;                          \
      case 2: LN_B2T(bf16, float, bf16, Gv, Jv); break;                          \
      case 3: LN_B2T(bf16, float, float, Gv, Jv); break;                         \
      case 4: LN_B2T(float, bf16, bf16, Gv, Jv); break;                          \
      case 5: LN_B2T(float, bf16, float, Gv, Jv); break;                         \
      case 6: LN_B2T(float, float, bf16, Gv, Jv); break;                         \
      default: LN_B2T(float, float, float, Gv, Jv); break;                       \
    }                                                                            \
  } while (0)
    LN_GJ_DISPATCH(LN_B2);
#undef LN_B2
#undef LN_B2T
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("ln_bwd_v2");
    return 0;
  }
  CMX_REQUIRE(!dbias, "layernorm_bwd: dbias needs the vectorised path (C %% 8 == 0, leading dims %% 8 == 0)");
  int grid = cdiv(M, 8);
  if (grid > 148 * 8) grid = 148 * 8;
#define LN_B(TDY, TX, TDX)                                                                                              \
  ln_bwd_kernel<TDY, TX, TDX><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const bf16*)dy2, lddy2, (const TX*)x, ldx, mean, \
                                                    rstd, gamma, dres, lddres, (TDX*)dx, lddx, (bf16*)dx_bf, lddxbf, scale, \
                                                    rows_per_sample, dgamma, dbeta, M, C)
  const int key = dy_dtype * 4 + x_dtype * 2 + dx_dtype;
  switch (key) {
    case 0: LN_B(bf16, bf16, bf16); break;
    case 1: LN_B(bf16, bf16, float); break;
    case 2: LN_B(bf16, float, bf16); break;
    case 3: LN_B(bf16, float, float); break;
    case 4: LN_B(float, bf16, bf16); break;
    case 5: LN_B(float, bf16, float); break;
    case 6: LN_B(float, float, bf16); break;
    default: LN_B(float, float, float); break;
  }
#undef LN_B
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("ln_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Column statistics (sum, sum of squares) in double.  Block = 32 channel lanes x 8 row lanes.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colstats_kernel(const T* __restrict__ x, long ldx, double* __restrict__ sum,
                                                       double* __restrict__ sumsq, long M, int C, int rows_per_cta) {
  pdl_trigger();
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a = 0.f, b = 0.f;
  if (c < C)
    for (long r = r0 + ty; r < r1; r += 8) {
      const float v = ld1(x + r * ldx + c);
      a += v;
      b += v * v;
    }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
    double da = 0., db = 0.;
#pragma unroll
    for (int i = 0; i < 8; i++) { da += (double)s1[i][tx]; db += (double)s2[i][tx]; }
    atomicAdd(sum + c, da);
    atomicAdd(sumsq + c, db);
  }
}

CMX_API int cmx_colstats(const void* x, int x_dtype, int64_t ldx, double* sum, double* sumsq, int64_t M, int C, void* stream) {
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int rows_per_cta = 256;
  dim3 grid(cdiv(C, 32), cdiv(M, rows_per_cta));
  CMX_REQUIRE(grid.y <= 65535, "colstats: M too large");
  if (x_dtype == CMX_F32) colstats_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ldx, sum, sumsq, M, C, rows_per_cta);
  else colstats_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, ldx, sum, sumsq, M, C, rows_per_cta);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("colstats");
  return 0;
}

__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, long count, float eps, float momentum,
                                   float* running_mean, float* running_var, int64_t* nbt, float* mean, float* invstd, int C) {
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double m = sum[c] / (double)count;
  double var = sumsq[c] / (double)count - m * m;
  if (var < 0.) var = 0.;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

CMX_API int cmx_bn_finalize(const double* sum, const double* sumsq, int64_t count, float eps, float momentum,
                            float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                            int C, void* stream) {
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, count, eps, momentum, running_mean,
                                                                    running_var, num_batches_tracked, mean, invstd, C);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("bn_finalize");
  return 0;
}

__global__ void bn_eval_stats_kernel(const float* rm, const float* rv, float eps, float* mean, float* invstd, int C) {
  pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.0f / sqrtf(rv[c] + eps);
}
CMX_API int cmx_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean, float* invstd, int C,
                              void* stream) {
  bn_eval_stats_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, eps, mean, invstd, C);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("bn_eval_stats");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// BN apply (+ residual, ReLU, Dropout2d mask):  thread = 4 consecutive channels of one row
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TR, typename TY>
__global__ void __launch_bounds__(256) bn_apply_kernel(const TX* __restrict__ x, long ldx, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const TR* __restrict__ res, long ldr,
                                                       int relu, const float* __restrict__ mask, int rows_per_sample,
                                                       TY* __restrict__ y, long ldy, long M, int C) {
  pdl_trigger();
  const int c4 = C >> 2;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * c4) return;
  const long row = idx / c4;
  const int c = (int)(idx % c4) * 4;
  float v[4], mu[4], is[4], g[4], b[4];
  load4(x + row * ldx + c, v);
  load4(mean + c, mu);
  load4(invstd + c, is);
  load4(gamma + c, g);
  load4(beta + c, b);
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = (v[i] - mu[i]) * is[i] * g[i] + b[i];
  if (res) {
    float r[4];
    load4(res + row * ldr + c, r);
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] += r[i];
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = fmaxf(v[i], 0.f);
  }
  if (mask) {
    float mk[4];
    load4(mask + (row / rows_per_sample) * C + c, mk);
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] *= mk[i];
  }
  store4(y + row * ldy + c, v);
}

template <typename TX, typename TR>
static void bn_apply_launch2(int y_dtype, const TX* x, long ldx, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, const TR* res, long ldr, int relu, const float* mask, int rps, void* y, long ldy,
                             long M, int C, cudaStream_t st) {
  const long n = M * (C >> 2);
  dim3 grid(cdiv(n, 256));
  if (y_dtype == CMX_F32)
    bn_apply_kernel<TX, TR, float><<<grid, 256, 0, st>>>(x, ldx, mean, invstd, gamma, beta, res, ldr, relu, mask, rps, (float*)y, ldy, M, C);
  else
    bn_apply_kernel<TX, TR, bf16><<<grid, 256, 0, st>>>(x, ldx, mean, invstd, gamma, beta, res, ldr, relu, mask, rps, (bf16*)y, ldy, M, C);
}

CMX_API int cmx_bn_apply(const void* x, int x_dtype, int64_t ldx, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const void* residual, int r_dtype, int64_t ldr, int relu, const float* mask,
                         int rows_per_sample, void* y, int y_dtype, int64_t ldy, int64_t M, int C, void* stream) {
  CMX_REQUIRE(C % 4 == 0, "bn_apply: C %% 4");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_per_sample <= 0) rows_per_sample = 1;
  if (x_dtype == CMX_F32) {
    if (r_dtype == CMX_F32) bn_apply_launch2<float, float>(y_dtype, (const float*)x, ldx, mean, invstd, gamma, beta, (const float*)residual, ldr, relu, mask, rows_per_sample, y, ldy, M, C, st);
    else bn_apply_launch2<float, bf16>(y_dtype, (const float*)x, ldx, mean, invstd, gamma, beta, (const bf16*)residual, ldr, relu, mask, rows_per_sample, y, ldy, M, C, st);
  } else {
    if (r_dtype == CMX_F32) bn_apply_launch2<bf16, float>(y_dtype, (const bf16*)x, ldx, mean, invstd, gamma, beta, (const float*)residual, ldr, relu, mask, rows_per_sample, y, ldy, M, C, st);
    else bn_apply_launch2<bf16, bf16>(y_dtype, (const bf16*)x, ldx, mean, invstd, gamma, beta, (const bf16*)residual, ldr, relu, mask, rows_per_sample, y, ldy, M, C, st);
  }
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("bn_apply");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// BN backward.  Effective upstream gradient g = dy * mask * [out > 0 if relu], out = bn(x)+res.
// pass 1: sum_g, sum_g_xhat per channel (double).  pass 2: dx = gamma*invstd*(g - sum_g/M - xhat*sum_g_xhat/M).
// ------------------------------------------------------------------------------------------------
template <typename TDY, typename TX, typename TR>
__device__ __forceinline__ float bn_eff_grad(const TDY* dy, long lddy, const TX* x, long ldx, const TR* res, long ldr, long row,
                                             int c, float mu, float is, float g, float b, int relu, const float* mask, int rps,
                                             int C, float& xhat) {
  const float xv = ld1(x + row * ldx + c);
  xhat = (xv - mu) * is;
  float d = ld1(dy + row * lddy + c);
  if (mask) d *= mask[(row / rps) * C + c];
  if (relu) {
    float o = xhat * g + b;
    if (res) o += ld1(res + row * ldr + c);
    if (o <= 0.f) d = 0.f;
  }
  return d;
}

template <typename TDY, typename TX, typename TR>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const TDY* __restrict__ dy, long lddy, const TX* __restrict__ x, long ldx,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const TR* __restrict__ res, long ldr, int relu,
                                                            const float* __restrict__ mask, int rps, double* __restrict__ sum_dy,
                                                            double* __restrict__ sum_dy_xhat, long M, int C, int rows_per_cta) {
  pdl_trigger();
  __shared__ float s1[8][33], s2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = mean[c], is = invstd[c], g = gamma[c], be = beta[c];
    for (long r = r0 + ty; r < r1; r += 8) {
      float xh;
      const float d = bn_eff_grad(dy, lddy, x, ldx, res, ldr, r, c, mu, is, g, be, relu, mask, rps, C, xh);
      a += d;
      b += d * xh;
    }
  }
  s1[ty][tx] = a;
  s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
    double da = 0., db = 0.;
#pragma unroll
    for (int i = 0; i < 8; i++) { da += (double)s1[i][tx]; db += (double)s2[i][tx]; }
    atomicAdd(sum_dy + c, da);
    atomicAdd(sum_dy_xhat + c, db);
  }
}

template <typename TDY, typename TX, typename TR, typename TDX>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const TDY* __restrict__ dy, long lddy, const TX* __restrict__ x, long ldx,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const TR* __restrict__ res, long ldr, int relu,
                                                           const float* __restrict__ mask, int rps,
                                                           const double* __restrict__ sum_dy, const double* __restrict__ sum_dy_xhat,
                                                           TDX* __restrict__ dx, long lddx, TDX* __restrict__ dres, long lddres,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int C) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C && blockIdx.y == 0 && dgamma) {
    // fold the parameter gradients into the first C threads of the grid
    atomicAdd(dgamma + idx, (float)sum_dy_xhat[idx]);
    atomicAdd(dbeta + idx, (float)sum_dy[idx]);
  }
  if (idx >= M * C) return;
  const long row = idx / C;
  const int c = (int)(idx % C);
  const float mu = mean[c], is = invstd[c], g = gamma[c], be = beta[c];
  float xh;
  const float d = bn_eff_grad(dy, lddy, x, ldx, res, ldr, row, c, mu, is, g, be, relu, mask, rps, C, xh);
  const float m1 = (float)(sum_dy[c] / (double)M), m2 = (float)(sum_dy_xhat[c] / (double)M);
  st1(dx + row * lddx + c, g * is * (d - m1 - xh * m2));
  if (dres) st1(dres + row * lddres + c, d);
}

// Vectorised variants (C % 8 == 0, C <= 2048, 16-byte aligned rows): thread = 8 consecutive channels (one or two
// 16-byte loads per tensor) of every NRL-th row, channel constants in registers, two rows in flight per thread.
template <typename TDY, typename TX, typename TR>
__device__ __forceinline__ void bn_eff_grad8(const TDY* dy, long lddy, const TX* x, long ldx, const TR* res, long ldr, long row,
                                             int c, const float* mu, const float* is, const float* g, const float* be, int relu,
                                             const float* mask, int rps, int C, float* d, float* xh) {
  float xv[8];
  load8(x + row * ldx + c, xv);
  load8(dy + row * lddy + c, d);
#pragma unroll
  for (int i = 0; i < 8; i++) xh[i] = (xv[i] - mu[i]) * is[i];
  if (mask) {
    float mk[8];
    load8(mask + (row / rps) * C + c, mk);
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] *= mk[i];
  }
  if (relu) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = xh[i] * g[i] + be[i];
    if (res) {
      float r[8];
      load8(res + row * ldr + c, r);
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] += r[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (o[i] <= 0.f) d[i] = 0.f;
  }
}

template <typename TDY, typename TX, typename TR>
__global__ void __launch_bounds__(256) bn_bwd_reduce_v8_kernel(const TDY* __restrict__ dy, long lddy, const TX* __restrict__ x,
                                                               long ldx, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const TR* __restrict__ res,
                                                               long ldr, int relu, const float* __restrict__ mask, int rps,
                                                               double* __restrict__ sum_dy, double* __restrict__ sum_dy_xhat,
                                                               long M, int C, int ng, int rows_per_cta) {
  pdl_trigger();
  __shared__ float sa[256][9], sb[256][9];
  const int tid = threadIdx.x;
  const int cg = tid % ng, rl = tid / ng, nrl = 256 / ng;
  const int c = (blockIdx.x * ng + cg) * 8;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = b[i] = 0.f;
  if (c < C && rl < nrl) {
    float mu[8], is[8], g[8], be[8];
    load8(mean + c, mu);
    load8(invstd + c, is);
    load8(gamma + c, g);
    load8(beta + c, be);
    long r = r0 + rl;
    for (; r + nrl < r1; r += 2 * nrl) {
      float d0[8], h0[8], d1[8], h1[8];
      bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r, c, mu, is, g, be, relu, mask, rps, C, d0, h0);
      bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r + nrl, c, mu, is, g, be, relu, mask, rps, C, d1, h1);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        a[i] += d0[i] + d1[i];
        b[i] += d0[i] * h0[i] + d1[i] * h1[i];
      }
    }
    if (r < r1) {
      float d0[8], h0[8];
      bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r, c, mu, is, g, be, relu, mask, rps, C, d0, h0);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        a[i] += d0[i];
        b[i] += d0[i] * h0[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sa[tid][i] = a[i];
    sb[tid][i] = b[i];
  }
  __syncthreads();
  if (tid < ng * 8) {
    const int gi = tid >> 3, i = tid & 7;
    const int cc = (blockIdx.x * ng + gi) * 8 + i;
    if (cc < C) {
      double da = 0., db = 0.;
      for (int l = 0; l < nrl; l++) {
        da += (double)sa[l * ng + gi][i];
        db += (double)sb[l * ng + gi][i];
      }
      atomicAdd(sum_dy + cc, da);
      atomicAdd(sum_dy_xhat + cc, db);
    }
  }
}

template <typename TDY, typename TX, typename TR, typename TDX>
__global__ void __launch_bounds__(256) bn_bwd_apply_v8_kernel(const TDY* __restrict__ dy, long lddy, const TX* __restrict__ x,
                                                              long ldx, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const TR* __restrict__ res,
                                                              long ldr, int relu, const float* __restrict__ mask, int rps,
                                                              const double* __restrict__ sum_dy,
                                                              const double* __restrict__ sum_dy_xhat, TDX* __restrict__ dx,
                                                              long lddx, TDX* __restrict__ dres, long lddres,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta, long M, int C,
                                                              int ng, int rows_per_cta) {
  pdl_trigger();
  const int tid = threadIdx.x;
  const int cg = tid % ng, rl = tid / ng, nrl = 256 / ng;
  const int c = (blockIdx.x * ng + cg) * 8;
  if (c >= C || rl >= nrl) return;
  const long r0 = (long)blockIdx.y * rows_per_cta;
  long r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float mu[8], is[8], g[8], be[8], m1[8], m2[8], gi[8];
  load8(mean + c, mu);
  load8(invstd + c, is);
  load8(gamma + c, g);
  load8(beta + c, be);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double s1 = sum_dy[c + i], s2 = sum_dy_xhat[c + i];
    m1[i] = (float)(s1 / (double)M);
    m2[i] = (float)(s2 / (double)M);
    gi[i] = g[i] * is[i];
    if (dgamma && blockIdx.y == 0 && rl == 0) {  // parameter gradients: once per channel
      atomicAdd(dgamma + c + i, (float)s2);
      atomicAdd(dbeta + c + i, (float)s1);
    }
  }
  long r = r0 + rl;
  for (; r + nrl < r1; r += 2 * nrl) {
    float d0[8], h0[8], d1[8], h1[8], o0[8], o1[8];
    bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r, c, mu, is, g, be, relu, mask, rps, C, d0, h0);
    bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r + nrl, c, mu, is, g, be, relu, mask, rps, C, d1, h1);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      o0[i] = gi[i] * (d0[i] - m1[i] - h0[i] * m2[i]);
      o1[i] = gi[i] * (d1[i] - m1[i] - h1[i] * m2[i]);
    }
    store8(dx + r * lddx + c, o0);
    store8(dx + (r + nrl) * lddx + c, o1);
    if (dres) {
      store8(dres + r * lddres + c, d0);
      store8(dres + (r + nrl) * lddres + c, d1);
    }
  }
  if (r < r1) {
    float d0[8], h0[8], o0[8];
    bn_eff_grad8(dy, lddy, x, ldx, res, ldr, r, c, mu, is, g, be, relu, mask, rps, C, d0, h0);
#pragma unroll
    for (int i = 0; i < 8; i++) o0[i] = gi[i] * (d0[i] - m1[i] - h0[i] * m2[i]);
    store8(dx + r * lddx + c, o0);
    if (dres) store8(dres + r * lddres + c, d0);
  }
}

// column-group / row-lane split of a 256-thread block and the row chunk that gives ~want_ctas CTAs
static inline void bn_v8_shape(long M, int C, long want_ctas, int& ng, int& gx, int& rows_per_cta) {
  ng = 32;  // largest power of two <= 32 that divides the number of 8-channel groups
  while ((C / 8) % ng) ng >>= 1;
  const int nrl = 256 / ng;
  gx = (C / 8) / ng;
  long rpc = (M * gx + want_ctas - 1) / want_ctas;
  rpc = (rpc + 2 * nrl - 1) / (2 * nrl) * (2 * nrl);
  if (rpc < 4L * nrl) rpc = 4L * nrl;
  rows_per_cta = (int)rpc;
}
static inline bool bn_v8_ok(int C, std::initializer_list<const void*> ptrs, std::initializer_list<long> lds) {
  if (C % 32) return false;  // at least four 16-byte groups per row segment (coalescing); otherwise the scalar kernels
  for (const void* p : ptrs)
    if (((uintptr_t)p) & 15) return false;
  for (long l : lds)
    if (l % 8) return false;
  return true;
}

#define BN_DISPATCH3(MACRO)                                                          \
  do {                                                                               \
    const int key = dy_dtype * 4 + x_dtype * 2 + r_dtype;                            \
    switch (key) {                                                                   \
      case 0: MACRO(bf16, bf16, bf16); break;                                        \
      case 1: MACRO(bf16, bf16, float); break;                                       \
      case 2: MACRO(bf16, float, bf16); break;                                       \
      case 3: MACRO(bf16, float, float); break;                                      \
      case 4: MACRO(float, bf16, bf16); break;                                       \
      case 5: MACRO(float, bf16, float); break;                                      \
      case 6: MACRO(float, float, bf16); break;                                      \
      default: MACRO(float, float, float); break;                                    \
    }                                                                                \
  } while (0)

CMX_API int cmx_bn_bwd_reduce(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                              const float* mean, const float* invstd, const float* gamma, const float* beta,
                              const void* residual, int r_dtype, int64_t ldr, int relu, const float* mask, int rows_per_sample,
                              double* sum_dy, double* sum_dy_xhat, int64_t M, int C, void* stream) {
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_per_sample <= 0) rows_per_sample = 1;
  if (bn_v8_ok(C, {dy, x, residual, mask, mean, invstd, gamma, beta}, {(long)lddy, (long)ldx, (long)ldr})) {
    int ng, gx, rpc;
    bn_v8_shape(M, C, 148L * 6, ng, gx, rpc);
    dim3 grid(gx, cdiv(M, rpc));
#define BN_R8(TDY, TX, TR)                                                                                                     \
  bn_bwd_reduce_v8_kernel<TDY, TX, TR><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma, beta, \
                                                             (const TR*)residual, ldr, relu, mask, rows_per_sample, sum_dy,     \
                                                             sum_dy_xhat, M, C, ng, rpc)
    BN_DISPATCH3(BN_R8);
#undef BN_R8
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("bn_bwd_reduce_v8");
    return 0;
  }
  const int rows_per_cta = 256;
  dim3 grid(cdiv(C, 32), cdiv(M, rows_per_cta));
  CMX_REQUIRE(grid.y <= 65535, "bn_bwd_reduce: M too large");
#define BN_R(TDY, TX, TR)                                                                                                   \
  bn_bwd_reduce_kernel<TDY, TX, TR><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma, beta, \
                                                          (const TR*)residual, ldr, relu, mask, rows_per_sample, sum_dy,     \
                                                          sum_dy_xhat, M, C, rows_per_cta)
  BN_DISPATCH3(BN_R);
#undef BN_R
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("bn_bwd_reduce");
  return 0;
}

CMX_API int cmx_bn_bwd_apply(const void* dy, int dy_dtype, int64_t lddy, const void* x, int x_dtype, int64_t ldx,
                             const float* mean, const float* invstd, const float* gamma, const float* beta, const void* residual,
                             int r_dtype, int64_t ldr, int relu, const float* mask, int rows_per_sample, const double* sum_dy,
                             const double* sum_dy_xhat, void* dx, int dx_dtype, int64_t lddx, void* dres, int dres_dtype,
                             int64_t lddres, float* dgamma, float* dbeta, int64_t M, int C, void* stream) {
  if (M == 0) return 0;
  CMX_REQUIRE(!dres || dres_dtype == dx_dtype, "bn_bwd_apply: dres dtype must equal dx dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows_per_sample <= 0) rows_per_sample = 1;
  if (bn_v8_ok(C, {dy, x, residual, mask, mean, invstd, gamma, beta, dx, dres},
               {(long)lddy, (long)ldx, (long)ldr, (long)lddx, (long)lddres})) {
    int ng, gx, rpc;
    bn_v8_shape(M, C, 148L * 8, ng, gx, rpc);
    dim3 grid8(gx, cdiv(M, rpc));
#define BN_A8(TDY, TX, TR)                                                                                                       \
  do {                                                                                                                           \
    if (dx_dtype == CMX_F32)                                                                                                     \
      bn_bwd_apply_v8_kernel<TDY, TX, TR, float><<<grid8, 256, 0, st>>>(                                                         \
          (const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma, beta, (const TR*)residual, ldr, relu, mask,              \
          rows_per_sample, sum_dy, sum_dy_xhat, (float*)dx, lddx, (float*)dres, lddres, dgamma, dbeta, M, C, ng, rpc);           \
    else                                                                                                                         \
      bn_bwd_apply_v8_kernel<TDY, TX, TR, bf16><<<grid8, 256, 0, st>>>(                                                          \
          (const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma, beta, (const TR*)residual, ldr, relu, mask,              \
          rows_per_sample, sum_dy, sum_dy_xhat, (bf16*)dx, lddx, (bf16*)dres, lddres, dgamma, dbeta, M, C, ng, rpc);             \
  } while (0)
    BN_DISPATCH3(BN_A8);
#undef BN_A8
    g_cmx_launches++;
    CMX_CHECK_LAUNCH("bn_bwd_apply_v8");
    return 0;
  }
  dim3 grid(cdiv(M * C, 256));
#define BN_A(TDY, TX, TR)                                                                                                       \
  do {                                                                                                                          \
    if (dx_dtype == CMX_F32)                                                                                                    \
      bn_bwd_apply_kernel<TDY, TX, TR, float><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma, \
                                                                    beta, (const TR*)residual, ldr, relu, mask, rows_per_sample,  \
                                                                    sum_dy, sum_dy_xhat, (float*)dx, lddx, (float*)dres, lddres,  \
                                                                    dgamma, dbeta, M, C);                                         \
    else                                                                                                                        \
      bn_bwd_apply_kernel<TDY, TX, TR, bf16><<<grid, 256, 0, st>>>((const TDY*)dy, lddy, (const TX*)x, ldx, mean, invstd, gamma,  \
                                                                   beta, (const TR*)residual, ldr, relu, mask, rows_per_sample,   \
                                                                   sum_dy, sum_dy_xhat, (bf16*)dx, lddx, (bf16*)dres, lddres,     \
                                                                   dgamma, dbeta, M, C);                                          \
  } while (0)
  BN_DISPATCH3(BN_A);
#undef BN_A
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("bn_bwd_apply");
  return 0;
}
