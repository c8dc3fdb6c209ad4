// Feature Rectify Module kernels (net_utils.py:22-30, 79-83, 147-152): global avg+max pooling with
// warp-shuffle / shared-memory reductions, the tiny channel-weight MLP (fp32, small-M), and the fused
// spatial-gate + rectification pass (one warp per token row) with its backward.
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

// ---- avg + max pool over the HW rows of each sample -------------------------------------------------------
// Stage 1: grid (column blocks, B, row chunks); thread = 8 channels (one 16-byte load) x row lane; per-chunk partial
// (sum, max, first-argmax) go to the caller's workspace.  Stage 2 folds the chunks in row order, so the FIRST maximum
// wins on ties (PyTorch adaptive_max_pool2d semantics).
constexpr int POOL_MAXCH = 64;
__global__ void __launch_bounds__(256) pool_partial_kernel(const bf16* __restrict__ x, long ldx, float* __restrict__ wsum,
                                                           float* __restrict__ wmax, int* __restrict__ widx, int HW, int C2,
                                                           int ng, int rows_per_chunk, int nchunk) {
  pdl_trigger();
  __shared__ float ssum[256][8];
  __shared__ float smax[256][8];
  __shared__ int sidx[256][8];
  const int tid = threadIdx.x;
  const int cg = tid % ng, rl = tid / ng, nrl = 256 / ng;
  const int c = (blockIdx.x * ng + cg) * 8;
  const int b = blockIdx.y, ch = blockIdx.z;
  const int r0 = ch * rows_per_chunk;
  int r1 = r0 + rows_per_chunk;
  if (r1 > HW) r1 = HW;
  float s[8], m[8];
  int mi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { s[i] = 0.f; m[i] = -INFINITY; mi[i] = r0; }
  if (c < C2 && rl < nrl)
    for (int r = r0 + rl; r < r1; r += nrl) {
      float v[8];
      load8(x + ((long)b * HW + r) * ldx + c, v);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        s[i] += v[i];
        if (v[i] > m[i]) { m[i] = v[i]; mi[i] = r; }
      }
    }
#pragma unroll
  for (int i = 0; i < 8; i++) { ssum[tid][i] = s[i]; smax[tid][i] = m[i]; sidx[tid][i] = mi[i]; }
  __syncthreads();
  if (tid < ng * 8) {
    const int g = tid >> 3, i = tid & 7;
    const int cc = (blockIdx.x * ng + g) * 8 + i;
    if (cc < C2) {
      float ts = 0.f, tm = -INFINITY;
      int ti = r0;
      for (int l = 0; l < nrl; l++) {
        const int t = l * ng + g;
        ts += ssum[t][i];
        const float v = smax[t][i];
        const int id = sidx[t][i];
        if (v > tm || (v == tm && id < ti)) { tm = v; ti = id; }
      }
      const long o = ((long)b * nchunk + ch) * C2 + cc;
      wsum[o] = ts; wmax[o] = tm; widx[o] = ti;
    }
  }
}
__global__ void pool_finalize_kernel(const float* __restrict__ wsum, const float* __restrict__ wmax, const int* __restrict__ widx,
                                     float* __restrict__ y, int32_t* __restrict__ argmax, int B, int HW, int C2, int nchunk) {
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C2) return;
  const int c = idx % C2, b = idx / C2;
  float ts = 0.f, tm = -INFINITY;
  int ti = 0;
  for (int ch = 0; ch < nchunk; ch++) {
    const long o = ((long)b * nchunk + ch) * C2 + c;
    ts += wsum[o];
    if (wmax[o] > tm) { tm = wmax[o]; ti = widx[o]; }
  }
  y[(long)b * 2 * C2 + c] = ts / (float)HW;
  y[(long)b * 2 * C2 + C2 + c] = tm;
  argmax[(long)b * C2 + c] = ti;
}
CMX_API int64_t cmx_pool_avgmax_ws_bytes(int B, int C2) { return (int64_t)B * POOL_MAXCH * C2 * 12; }
CMX_API int cmx_pool_avgmax_fwd(const void* x, int64_t ldx, float* y, int32_t* argmax, void* ws, int B, int HW, int C2, void* stream) {
  if (B == 0) return 0;
  CMX_REQUIRE(C2 % 8 == 0 && ldx % 8 == 0 && ws, "pool_avgmax: C2 %% 8 and a workspace are required");
  cudaStream_t st = (cudaStream_t)stream;
  int ng = C2 / 8;
  if (ng > 32) ng = 32;
  while (256 % ng) ng--;
  const int nrl = 256 / ng;
  int nchunk = cdiv(HW, 4 * nrl);
  if (nchunk > POOL_MAXCH) nchunk = POOL_MAXCH;
  if (nchunk < 1) nchunk = 1;
  const int rows_per_chunk = cdiv(HW, nchunk);
  nchunk = cdiv(HW, rows_per_chunk);
  float* wsum = (float*)ws;
  float* wmax = wsum + (size_t)B * POOL_MAXCH * C2;
  int* widx = (int*)(wmax + (size_t)B * POOL_MAXCH * C2);
  dim3 grid(cdiv(C2 / 8, ng), B, nchunk);
  pool_partial_kernel<<<grid, 256, 0, st>>>((const bf16*)x, ldx, wsum, wmax, widx, HW, C2, ng, rows_per_chunk, nchunk);
  g_cmx_launches++;
  pool_finalize_kernel<<<cdiv(B * C2, 128), 128, 0, st>>>(wsum, wmax, widx, y, argmax, B, HW, C2, nchunk);
  LAUNCH_DONE("pool_avgmax_fwd");
}
__global__ void __launch_bounds__(256) pool_avgmax_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ argmax,
                                                              float* __restrict__ dx, long lddx, int HW, int C2, long total) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C2);
  const long row = idx / C2;
  const int b = (int)(row / HW);
  const int r = (int)(row % HW);
  float g = dy[(long)b * 2 * C2 + c] / (float)HW;
  if (argmax[(long)b * C2 + c] == r) g += dy[(long)b * 2 * C2 + C2 + c];
  dx[row * lddx + c] += g;
}
CMX_API int cmx_pool_avgmax_bwd(const float* dy, const int32_t* argmax, float* dx, int64_t lddx, int B, int HW, int C2, void* stream) {
  const long total = (long)B * HW * C2;
  if (total == 0) return 0;
  pool_avgmax_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, argmax, dx, lddx, HW, C2, total);
  LAUNCH_DONE("pool_avgmax_bwd");
}

// ---- small-M fp32 linear: one warp per output column, all Mb <= 16 rows at once ------------------------
constexpr int SMALLM_MAX = 16;
__global__ void __launch_bounds__(256) smallm_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ b, int act, float* __restrict__ y, int Mb,
                                                                int N, int K) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[SMALLM_MAX];
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++) acc[m] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[(long)n * K + k];
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++)
      if (m < Mb) acc[m] = fmaf(wv, x[(long)m * K + k], acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++) {
    if (m < Mb) {
      float v = warp_sum(acc[m]);
      if (lane == 0) {
        if (b) v += b[n];
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 3) v = sigmoid_f(v);
        y[(long)m * N + n] = v;
      }
    }
  }
}
CMX_API int cmx_smallm_linear_fwd(const float* x, const float* w, const float* b, int act, float* y, int Mb, int N, int K, void* stream) {
  CMX_REQUIRE(Mb >= 1 && Mb <= SMALLM_MAX, "smallm_linear: Mb=%d must be in [1,%d]", Mb, SMALLM_MAX);
  smallm_linear_fwd_kernel<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(x, w, b, act, y, Mb, N, K);
  LAUNCH_DONE("smallm_linear_fwd");
}
// dpre = dy * act'(y)
__global__ void smallm_dpre_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act, float* __restrict__ dpre,
                                   float* __restrict__ db, int Mb, int N) {
  pdl_trigger();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < Mb; m++) {
    const float yv = y[(long)m * N + n];
    float g = dy[(long)m * N + n];
    if (act == 1) g = yv > 0.f ? g : 0.f;
    else if (act == 3) g *= yv * (1.f - yv);
    dpre[(long)m * N + n] = g;
    s += g;
  }
  if (db) db[n] += s;
}
// dW[n,k] += sum_m dpre[m,n] x[m,k];   dx[m,k] = sum_n dpre[m,n] W[n,k]
__global__ void __launch_bounds__(256) smallm_dw_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dw,
                                                        int Mb, int N, int K) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * K) return;
  const int k = (int)(idx % K);
  const int n = (int)(idx / K);
  float s = 0.f;
  for (int m = 0; m < Mb; m++) s = fmaf(dpre[(long)m * N + n], x[(long)m * K + k], s);
  dw[idx] += s;
}
// dx[m,k] = sum_n dpre[m,n] W[n,k]: thread = column k (coalesced W reads), blockIdx.y = chunk of 64 rows of W,
// all Mb <= 16 samples at once; partial sums are atomically added into the zeroed dx
__global__ void __launch_bounds__(256) smallm_dx_kernel(const float* __restrict__ dpre, const float* __restrict__ w, float* __restrict__ dx,
                                                        int Mb, int N, int K) {
  pdl_trigger();
  __shared__ float sd[SMALLM_MAX][64];
  const int n0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < SMALLM_MAX * 64; i += 256) {
    const int m = i >> 6, n = n0 + (i & 63);
    sd[m][i & 63] = (m < Mb && n < N) ? dpre[(long)m * N + n] : 0.f;
  }
  __syncthreads();
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  float acc[SMALLM_MAX];
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++) acc[m] = 0.f;
  const int nn = min(64, N - n0);
  for (int n = 0; n < nn; n++) {
    const float wv = w[(long)(n0 + n) * K + k];
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++) acc[m] = fmaf(sd[m][n], wv, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++)
    if (m < Mb) atomicAdd(dx + (long)m * K + k, acc[m]);
}
CMX_API int cmx_smallm_linear_bwd(const float* dy, const float* y, int act, const float* x, const float* w, float* dx, float* dw,
                                  float* db, float* dpre_ws, int Mb, int N, int K, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  smallm_dpre_kernel<<<cdiv(N, 128), 128, 0, st>>>(dy, y, act, dpre_ws, db, Mb, N);
  g_cmx_launches++;
  if (dw) {
    smallm_dw_kernel<<<cdiv((long)N * K, 256), 256, 0, st>>>(dpre_ws, x, dw, Mb, N, K);
    g_cmx_launches++;
  }
  if (dx) {
    cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)Mb * K, st);
    dim3 grid(cdiv(K, 256), cdiv(N, 64));
    smallm_dx_kernel<<<grid, 256, 0, st>>>(dpre_ws, w, dx, Mb, N, K);
    g_cmx_launches++;
  }
  CMX_CHECK_LAUNCH("smallm_linear_bwd");
  return 0;
}

// ---- fused spatial gate + rectification ------------------------------------------------------------------
// One warp per token row; lane owns channels {4*(lane+32j)}, C <= 512, C % 4 == 0.
constexpr int FR_MAXJ = 4;
__global__ void __launch_bounds__(256) frm_rectify_fwd_kernel(const bf16* __restrict__ a, long lda, const bf16* __restrict__ t, long ldt,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              const float* __restrict__ cw, float* __restrict__ sw,
                                                              bf16* __restrict__ r1, long ldr1, bf16* __restrict__ r2, long ldr2,
                                                              long M, int HW, int C) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int b = (int)(row / HW);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < FR_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
      float tv[4], wa[4], wb[4];
      load4(t + row * ldt + c, tv);
      load4(w2 + c, wa);
      load4(w2 + C + c, wb);
#pragma unroll
      for (int i = 0; i < 4; i++) { d0 = fmaf(tv[i], wa[i], d0); d1 = fmaf(tv[i], wb[i], d1); }
    }
  }
  const float s0 = sigmoid_f(warp_sum(d0) + b2[0]);
  const float s1 = sigmoid_f(warp_sum(d1) + b2[1]);
  if (lane == 0) { sw[row * 2] = s0; sw[row * 2 + 1] = s1; }
#pragma unroll
  for (int j = 0; j < FR_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
      float a1[4], a2[4], c0[4], c1[4], o1[4], o2[4];
      load4(a + row * lda + c, a1);
      load4(a + row * lda + C + c, a2);
      load4(cw + (long)b * 2 * C + c, c0);
      load4(cw + (long)b * 2 * C + C + c, c1);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        o1[i] = a1[i] + 0.5f * c1[i] * a2[i] + 0.5f * s1 * a2[i];
        o2[i] = a2[i] + 0.5f * c0[i] * a1[i] + 0.5f * s0 * a1[i];
      }
      store4(r1 + row * ldr1 + c, o1);
      store4(r2 + row * ldr2 + c, o2);
    }
  }
}
CMX_API int cmx_frm_rectify_fwd(const void* a, int64_t lda, const void* t, int64_t ldt, const float* w2, const float* b2,
                                const float* cw, float* sw, void* r1, int64_t ldr1, void* r2, int64_t ldr2, int B, int HW, int C,
                                void* stream) {
  CMX_REQUIRE(C % 4 == 0 && C <= 128 * FR_MAXJ, "frm_rectify: C=%d unsupported", C);
  const long M = (long)B * HW;
  if (M == 0) return 0;
  frm_rectify_fwd_kernel<<<cdiv(M, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, lda, (const bf16*)t, ldt, w2, b2, cw, sw,
                                                                       (bf16*)r1, ldr1, (bf16*)r2, ldr2, M, HW, C);
  LAUNCH_DONE("frm_rectify_fwd");
}

// backward.  grid (ctas_per_sample, B): every CTA stays inside one sample so the per-(b,c) channel-weight
// gradients can be reduced in shared memory before one atomicAdd per channel per CTA.
__global__ void __launch_bounds__(256) frm_rectify_bwd_kernel(const float* __restrict__ dr1, long lddr1, const float* __restrict__ dr2,
                                                              long lddr2, const bf16* __restrict__ a, long lda,
                                                              const bf16* __restrict__ t, long ldt, const float* __restrict__ w2,
                                                              const float* __restrict__ cw, const float* __restrict__ sw,
                                                              float* __restrict__ da, long ldda, bf16* __restrict__ dt, long lddt,
                                                              float* __restrict__ dcw, float* __restrict__ dw2, float* __restrict__ db2,
                                                              int HW, int C) {
  pdl_trigger();
  __shared__ float s_dcw[2][512];
  __shared__ float s_dw2[2][512];
  __shared__ float s_db2[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * 512; i += blockDim.x) { (&s_dcw[0][0])[i] = 0.f; (&s_dw2[0][0])[i] = 0.f; }
  if (threadIdx.x < 2) s_db2[threadIdx.x] = 0.f;
  __syncthreads();
  float acw0[FR_MAXJ][4], acw1[FR_MAXJ][4], aw0[FR_MAXJ][4], aw1[FR_MAXJ][4];
#pragma unroll
  for (int j = 0; j < FR_MAXJ; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) { acw0[j][i] = 0.f; acw1[j][i] = 0.f; aw0[j][i] = 0.f; aw1[j][i] = 0.f; }
  float ab0 = 0.f, ab1 = 0.f;
  for (int r = blockIdx.x * nwarp + warp; r < HW; r += gridDim.x * nwarp) {
    const long row = (long)b * HW + r;
    const float s0 = sw[row * 2], s1 = sw[row * 2 + 1];
    float g1[FR_MAXJ][4], g2[FR_MAXJ][4], a1[FR_MAXJ][4], a2[FR_MAXJ][4];
    float ds0 = 0.f, ds1 = 0.f;
#pragma unroll
    for (int j = 0; j < FR_MAXJ; j++) {
      const int c = 4 * (lane + 32 * j);
      if (c < C) {
        load4(dr1 + row * lddr1 + c, g1[j]);
        load4(dr2 + row * lddr2 + c, g2[j]);
        load4(a + row * lda + c, a1[j]);
        load4(a + row * lda + C + c, a2[j]);
        float c0[4], c1[4], o1[4], o2[4];
        load4(cw + (long)b * 2 * C + c, c0);
        load4(cw + (long)b * 2 * C + C + c, c1);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float p1 = g1[j][i] * a2[j][i];  // d/d(gate1) contributions
          const float p0 = g2[j][i] * a1[j][i];
          ds1 += p1; ds0 += p0;
          acw1[j][i] += 0.5f * p1;
          acw0[j][i] += 0.5f * p0;
          o1[i] = g1[j][i] + 0.5f * (c0[i] + s0) * g2[j][i];
          o2[i] = g2[j][i] + 0.5f * (c1[i] + s1) * g1[j][i];
        }
        store4(da + row * ldda + c, o1);
        store4(da + row * ldda + C + c, o2);
      }
    }
    ds0 = 0.5f * warp_sum(ds0) * s0 * (1.f - s0);  // through the sigmoid
    ds1 = 0.5f * warp_sum(ds1) * s1 * (1.f - s1);
    ab0 += ds0; ab1 += ds1;
#pragma unroll
    for (int j = 0; j < FR_MAXJ; j++) {
      const int c = 4 * (lane + 32 * j);
      if (c < C) {
        float tv[4], wa[4], wb[4], o[4];
        load4(t + row * ldt + c, tv);
        load4(w2 + c, wa);
        load4(w2 + C + c, wb);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          aw0[j][i] = fmaf(ds0, tv[i], aw0[j][i]);
          aw1[j][i] = fmaf(ds1, tv[i], aw1[j][i]);
          o[i] = tv[i] > 0.f ? (ds0 * wa[i] + ds1 * wb[i]) : 0.f;  // ReLU of the hidden 1x1 conv
        }
        store4(dt + row * lddt + c, o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < FR_MAXJ; j++) {
    const int c = 4 * (lane + 32 * j);
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        atomicAdd(&s_dcw[0][c + i], acw0[j][i]);
        atomicAdd(&s_dcw[1][c + i], acw1[j][i]);
        atomicAdd(&s_dw2[0][c + i], aw0[j][i]);
        atomicAdd(&s_dw2[1][c + i], aw1[j][i]);
      }
    }
  }
  if (lane == 0) { atomicAdd(&s_db2[0], ab0); atomicAdd(&s_db2[1], ab1); }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dcw + (long)b * 2 * C + i, s_dcw[0][i]);
    atomicAdd(dcw + (long)b * 2 * C + C + i, s_dcw[1][i]);
    atomicAdd(dw2 + i, s_dw2[0][i]);
    atomicAdd(dw2 + C + i, s_dw2[1][i]);
  }
  if (threadIdx.x < 2) atomicAdd(db2 + threadIdx.x, s_db2[threadIdx.x]);
}
CMX_API int cmx_frm_rectify_bwd(const float* dr1, int64_t lddr1, const float* dr2, int64_t lddr2, const void* a, int64_t lda,
                                const void* t, int64_t ldt, const float* w2, const float* cw, const float* sw, float* da,
                                int64_t ldda, void* dt, int64_t lddt, float* dcw, float* dw2, float* db2, int B, int HW, int C,
                                void* stream) {
  CMX_REQUIRE(C % 4 == 0 && C <= 128 * FR_MAXJ, "frm_rectify_bwd: C=%d unsupported", C);
  if (B == 0 || HW == 0) return 0;
  int gx = cdiv(HW, 8 * 8);  // >= 8 rows per warp before the reduction flush
  if (gx < 1) gx = 1;
  if (gx > 296) gx = 296;
  dim3 grid(gx, B);
  frm_rectify_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dr1, lddr1, dr2, lddr2, (const bf16*)a, lda, (const bf16*)t, ldt, w2,
                                                                 cw, sw, da, ldda, (bf16*)dt, lddt, dcw, dw2, db2, HW, C);
  LAUNCH_DONE("frm_rectify_bwd");
}
