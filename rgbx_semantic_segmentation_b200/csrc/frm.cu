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
// block = 32 channels x 8 chunk lanes: the (up to 64) chunk partials of a channel are read by 8 lanes in parallel (this
// kernel is pure load latency) and folded in shared memory; ties go to the smaller row index (= first maximum)
__global__ void __launch_bounds__(256) pool_finalize_kernel(const float* __restrict__ wsum, const float* __restrict__ wmax,
                                                            const int* __restrict__ widx, float* __restrict__ y,
                                                            int32_t* __restrict__ argmax, int B, int HW, int C2, int nchunk) {
  pdl_trigger();
  __shared__ float ssum[8][33], smax[8][33];
  __shared__ int sidx[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + tx;
  const bool ok = idx < B * C2;
  const int c = ok ? idx % C2 : 0, b = ok ? idx / C2 : 0;
  float ts = 0.f, tm = -INFINITY;
  int ti = 0x7fffffff;
  if (ok) {
#pragma unroll 8
    for (int ch = ty; ch < nchunk; ch += 8) {
      const long o = ((long)b * nchunk + ch) * C2 + c;
      const float v = wmax[o];
      ts += wsum[o];
      if (v > tm) { tm = v; ti = widx[o]; }
    }
  }
  ssum[ty][tx] = ts; smax[ty][tx] = tm; sidx[ty][tx] = ti;
  __syncthreads();
  if (ty == 0 && ok) {
#pragma unroll
    for (int l = 1; l < 8; l++) {
      ts += ssum[l][tx];
      const float v = smax[l][tx];
      const int id = sidx[l][tx];
      if (v > tm || (v == tm && id < ti)) { tm = v; ti = id; }
    }
    y[(long)b * 2 * C2 + c] = ts / (float)HW;
    y[(long)b * 2 * C2 + C2 + c] = tm;
    argmax[(long)b * C2 + c] = ti == 0x7fffffff ? 0 : ti;
  }
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
  pool_finalize_kernel<<<cdiv(B * C2, 32), 256, 0, st>>>(wsum, wmax, widx, y, argmax, B, HW, C2, nchunk);
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
// thread = 4 consecutive channels of one row (16-byte read-modify-write)
__global__ void __launch_bounds__(256) pool_avgmax_bwd_v4_kernel(const float* __restrict__ dy, const int32_t* __restrict__ argmax,
                                                                 float* __restrict__ dx, long lddx, int HW, int C2, long total4) {
  pdl_trigger();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int c4 = C2 >> 2;
  const int c = (int)(idx % c4) * 4;
  const long row = idx / c4;
  const int b = (int)(row / HW);
  const int r = (int)(row - (long)b * HW);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(dy + (long)b * 2 * C2 + c));
  const float4 gm = __ldg(reinterpret_cast<const float4*>(dy + (long)b * 2 * C2 + C2 + c));
  const int4 am = __ldg(reinterpret_cast<const int4*>(argmax + (long)b * C2 + c));
  float4* p = reinterpret_cast<float4*>(dx + row * lddx + c);
  float4 v = *p;
  const float inv = 1.f / (float)HW;
  v.x += ga.x * inv + (am.x == r ? gm.x : 0.f);
  v.y += ga.y * inv + (am.y == r ? gm.y : 0.f);
  v.z += ga.z * inv + (am.z == r ? gm.z : 0.f);
  v.w += ga.w * inv + (am.w == r ? gm.w : 0.f);
  *p = v;
}
CMX_API int cmx_pool_avgmax_bwd(const float* dy, const int32_t* argmax, float* dx, int64_t lddx, int B, int HW, int C2, void* stream) {
  const long total = (long)B * HW * C2;
  if (total == 0) return 0;
  if (C2 % 4 == 0 && lddx % 4 == 0 && ((((uintptr_t)dy) | ((uintptr_t)argmax) | ((uintptr_t)dx)) & 15) == 0) {
    pool_avgmax_bwd_v4_kernel<<<cdiv(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(dy, argmax, dx, lddx, HW, C2, total / 4);
    LAUNCH_DONE("pool_avgmax_bwd");
  }
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
// K % 4 == 0: 16-byte loads and four independent weight loads in flight per lane (the scalar loop above serialises
// K/32 DRAM round trips per warp: 71 us for the 2048 x 2048 stage-4 matrix)
__global__ void __launch_bounds__(256) smallm_linear_fwd_v4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ b, int act, float* __restrict__ y,
                                                                   int Mb, int N, int K) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[SMALLM_MAX];
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++) acc[m] = 0.f;
  const float* wr = w + (long)n * K;
  int k = lane * 4;
  // eight, then four independent 16-byte weight loads in flight per lane (ncu: the kernel is a chain of DRAM round trips -
  // long-scoreboard stalls 10-20x the issue time - so the depth of each round trip is what counts)
  for (; k + 7 * 128 < K; k += 8 * 128) {
    float4 wv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) wv[u] = __ldg(reinterpret_cast<const float4*>(wr + k + u * 128));
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++)
      if (m < Mb) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long)m * K + k + u * 128));
          acc[m] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, acc[m]))));
        }
      }
  }
  for (; k + 3 * 128 < K; k += 4 * 128) {
    float4 wv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) wv[u] = __ldg(reinterpret_cast<const float4*>(wr + k + u * 128));
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++)
      if (m < Mb) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long)m * K + k + u * 128));
          acc[m] = fmaf(wv[u].x, xv.x, fmaf(wv[u].y, xv.y, fmaf(wv[u].z, xv.z, fmaf(wv[u].w, xv.w, acc[m]))));
        }
      }
  }
  for (; k < K; k += 128) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++)
      if (m < Mb) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long)m * K + k));
        acc[m] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[m]))));
      }
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
  if (K % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0)
    smallm_linear_fwd_v4_kernel<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(x, w, b, act, y, Mb, N, K);
  else
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
  // 8 independent weight loads in flight (rows beyond N are zero in sd, their weights are read as 0)
  for (int n = 0; n < 64; n += 8) {
    if (n0 + n >= N) break;
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; u++) wv[u] = n0 + n + u < N ? __ldg(w + (long)(n0 + n + u) * K + k) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int m = 0; m < SMALLM_MAX; m++) acc[m] = fmaf(sd[m][n + u], wv[u], acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++)
    if (m < Mb) atomicAdd(dx + (long)m * K + k, acc[m]);
}
// ONE launch for the whole backward of a small-M linear (the three kernels above stay for reference / CMX_SMALLM_BWD_SPLIT=1):
//   CTAs [0, nb_dw):      dW[n,k] += sum_m dpre[m,n] x[m,k], four consecutive k per thread (16-byte accesses)
//   CTAs [nb_dw, ...):    dx[m,k] += sum_{n in chunk} dpre[m,n] W[n,k] (atomics into the zeroed dx); the k-chunk-0 CTAs also add db
// with dpre = dy * act'(y) recomputed where it is used (Mb * N values) instead of a workspace round trip.
__device__ __forceinline__ float smallm_dpre(float g, float yv, int act) {
  if (act == 1) return yv > 0.f ? g : 0.f;
  if (act == 3) return g * yv * (1.f - yv);
  return g;
}
__global__ void __launch_bounds__(256) smallm_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                                               const float* __restrict__ x, const float* __restrict__ w,
                                                               float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                               int Mb, int N, int K, int nb_dw, int kchunks) {
  pdl_trigger();
  __shared__ float sd[SMALLM_MAX][64];
  if ((int)blockIdx.x < nb_dw) {
    // ---- dW: K % 4 == 0 (checked by the launcher); one thread = 4 consecutive k of one row n
    const long q = (long)blockIdx.x * 256 + threadIdx.x;
    const int kq = K >> 2;
    if (q >= (long)N * kq) return;
    const int n = (int)(q / kq), k = (int)(q % kq) * 4;
    float4 s = *reinterpret_cast<const float4*>(dw + (long)n * K + k);
    // all loads of the (predicated, fully unrolled) sample loop are issued before the first use: one memory round trip
    float gy[SMALLM_MAX], gv[SMALLM_MAX];
    float4 xv[SMALLM_MAX];
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++) {
      gy[m] = 0.f; gv[m] = 0.f; xv[m] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < Mb) {
        gy[m] = __ldg(dy + (long)m * N + n);
        gv[m] = __ldg(y + (long)m * N + n);
        xv[m] = __ldg(reinterpret_cast<const float4*>(x + (long)m * K + k));
      }
    }
#pragma unroll
    for (int m = 0; m < SMALLM_MAX; m++) {
      const float g = smallm_dpre(gy[m], gv[m], act);   // rows >= Mb: g = dpre(0, 0) = 0
      s.x = fmaf(g, xv[m].x, s.x); s.y = fmaf(g, xv[m].y, s.y); s.z = fmaf(g, xv[m].z, s.z); s.w = fmaf(g, xv[m].w, s.w);
    }
    *reinterpret_cast<float4*>(dw + (long)n * K + k) = s;
    return;
  }
  const int bid = (int)blockIdx.x - nb_dw;
  const int kc = bid % kchunks, n0 = (bid / kchunks) * 64;
  {
    float a_[4], b_[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {   // SMALLM_MAX * 64 / 256 = 4 entries per thread, loads first
      const int i = threadIdx.x + u * 256, m = i >> 6, n = n0 + (i & 63);
      const bool ok = m < Mb && n < N;
      a_[u] = ok ? __ldg(dy + (long)m * N + n) : 0.f;
      b_[u] = ok ? __ldg(y + (long)m * N + n) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = threadIdx.x + u * 256;
      sd[i >> 6][i & 63] = smallm_dpre(a_[u], b_[u], act);
    }
  }
  __syncthreads();
  if (db && kc == 0 && threadIdx.x < 64 && n0 + (int)threadIdx.x < N) {
    float sb = 0.f;
    for (int m = 0; m < Mb; m++) sb += sd[m][threadIdx.x];
    db[n0 + threadIdx.x] += sb;
  }
  if (!dx) return;
  const int k = kc * 256 + threadIdx.x;
  if (k >= K) return;
  float acc[SMALLM_MAX];
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++) acc[m] = 0.f;
  for (int n = 0; n < 64; n += 32) {   // 32 weight loads in flight per round (two rounds per 64-row chunk)
    if (n0 + n >= N) break;
    float wv[32];
#pragma unroll
    for (int u = 0; u < 32; u++) wv[u] = n0 + n + u < N ? __ldg(w + (long)(n0 + n + u) * K + k) : 0.f;
#pragma unroll
    for (int u = 0; u < 32; u++)
#pragma unroll
      for (int m = 0; m < SMALLM_MAX; m++) acc[m] = fmaf(sd[m][n + u], wv[u], acc[m]);
  }
#pragma unroll
  for (int m = 0; m < SMALLM_MAX; m++)
    if (m < Mb) atomicAdd(dx + (long)m * K + k, acc[m]);
}
CMX_API int cmx_smallm_linear_bwd(const float* dy, const float* y, int act, const float* x, const float* w, float* dx, float* dw,
                                  float* db, float* dpre_ws, int Mb, int N, int K, void* stream) {
  CMX_REQUIRE(Mb >= 1 && Mb <= SMALLM_MAX, "smallm_linear_bwd: Mb=%d must be in [1,%d]", Mb, SMALLM_MAX);
  cudaStream_t st = (cudaStream_t)stream;
  static int split = -1;
  if (split < 0) split = getenv("CMX_SMALLM_BWD_SPLIT") != nullptr ? 1 : 0;
  if (!split && K % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)(dw ? dw : x))) & 15) == 0) {
    if (dx) cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)Mb * K, st);
    const int nb_dw = dw ? (int)cdiv((long)N * (K / 4), 256) : 0;
    const int kchunks = (int)cdiv(K, 256);
    const int nb_dx = (dx || db) ? kchunks * (int)cdiv(N, 64) : 0;
    smallm_bwd_fused_kernel<<<nb_dw + nb_dx, 256, 0, st>>>(dy, y, act, x, w, dx, dw, db, Mb, N, K, nb_dw, kchunks);
    LAUNCH_DONE("smallm_linear_bwd");
  }
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
// sum over the G (power of two <= 32) consecutive lanes that share a row
__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// G lanes per row: 32, or C/4 when that is a smaller power of two (C = 64: two rows per warp, no idle lanes)
static inline int frm_lanes_per_row(int C) {
  const int g = C / 4;
  return (g < 32 && (g & (g - 1)) == 0) ? g : 32;
}
__global__ void __launch_bounds__(256) frm_rectify_fwd_kernel(const bf16* __restrict__ a, long lda, const bf16* __restrict__ t, long ldt,
                                                              const float* __restrict__ w2, const float* __restrict__ b2,
                                                              const float* __restrict__ cw, float* __restrict__ sw,
                                                              bf16* __restrict__ r1, long ldr1, bf16* __restrict__ r2, long ldr2,
                                                              long M, int HW, int C, int G) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int sl = lane % G, rpw = 32 / G;
  const long row = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + lane / G;
  const bool valid = row < M;
  const int b = valid ? (int)(row / HW) : 0;
  float d0 = 0.f, d1 = 0.f;
  if (valid) {
#pragma unroll
    for (int j = 0; j < FR_MAXJ; j++) {
      const int c = 4 * (sl + 32 * j);
      if (c < C) {
        float tv[4], wa[4], wb[4];
        load4(t + row * ldt + c, tv);
        load4(w2 + c, wa);
        load4(w2 + C + c, wb);
#pragma unroll
        for (int i = 0; i < 4; i++) { d0 = fmaf(tv[i], wa[i], d0); d1 = fmaf(tv[i], wb[i], d1); }
      }
    }
  }
  const float s0 = sigmoid_f(group_sum(d0, G) + b2[0]);
  const float s1 = sigmoid_f(group_sum(d1, G) + b2[1]);
  if (!valid) return;
  if (sl == 0) { sw[row * 2] = s0; sw[row * 2 + 1] = s1; }
#pragma unroll
  for (int j = 0; j < FR_MAXJ; j++) {
    const int c = 4 * (sl + 32 * j);
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
  const int G = frm_lanes_per_row(C);
  frm_rectify_fwd_kernel<<<cdiv(M, 8 * (32 / G)), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, lda, (const bf16*)t, ldt, w2, b2, cw,
                                                                                  sw, (bf16*)r1, ldr1, (bf16*)r2, ldr2, M, HW, C, G);
  LAUNCH_DONE("frm_rectify_fwd");
}

// backward.  grid (ctas_per_sample, B): every CTA stays inside one sample so the per-(b,c) channel-weight
// gradients can be reduced in shared memory before one atomicAdd per channel per CTA.
// J = ceil(C / 128) register slices per lane (compile time: the per-lane accumulators are 16 floats per slice)
template <int J>
__global__ void __launch_bounds__(256) frm_rectify_bwd_kernel(const float* __restrict__ dr1, long lddr1, const float* __restrict__ dr2,
                                                              long lddr2, const bf16* __restrict__ a, long lda,
                                                              const bf16* __restrict__ t, long ldt, const float* __restrict__ w2,
                                                              const float* __restrict__ cw, const float* __restrict__ sw,
                                                              float* __restrict__ da, long ldda, bf16* __restrict__ dt, long lddt,
                                                              float* __restrict__ dcw, float* __restrict__ dw2, float* __restrict__ db2,
                                                              int HW, int C, int G) {
  pdl_trigger();
  __shared__ float s_dcw[2][512];
  __shared__ float s_dw2[2][512];
  __shared__ float s_db2[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * 512; i += blockDim.x) { (&s_dcw[0][0])[i] = 0.f; (&s_dw2[0][0])[i] = 0.f; }
  if (threadIdx.x < 2) s_db2[threadIdx.x] = 0.f;
  __syncthreads();
  float acw0[J][4], acw1[J][4], aw0[J][4], aw1[J][4];
#pragma unroll
  for (int j = 0; j < J; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) { acw0[j][i] = 0.f; acw1[j][i] = 0.f; aw0[j][i] = 0.f; aw1[j][i] = 0.f; }
  float ab0 = 0.f, ab1 = 0.f;
  const int sl = lane % G, rpw = 32 / G, sub = lane / G;
  for (int rb = (blockIdx.x * nwarp + warp) * rpw; rb < HW; rb += gridDim.x * nwarp * rpw) {
    const int r = rb + sub;
    const bool valid = r < HW;
    const long row = (long)b * HW + (valid ? r : rb);  // invalid sub-rows shadow a valid one, contribute nothing, store nothing
    const float s0 = sw[row * 2], s1 = sw[row * 2 + 1];
    float g1[J][4], g2[J][4], a1[J][4], a2[J][4];
    float ds0 = 0.f, ds1 = 0.f;
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 4 * (sl + 32 * j);
      if (c < C && valid) {
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
    ds0 = 0.5f * group_sum(ds0, G) * s0 * (1.f - s0);  // through the sigmoid
    ds1 = 0.5f * group_sum(ds1, G) * s1 * (1.f - s1);
    if (sl == 0) { ab0 += ds0; ab1 += ds1; }
#pragma unroll
    for (int j = 0; j < J; j++) {
      const int c = 4 * (sl + 32 * j);
      if (c < C && valid) {
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
  for (int j = 0; j < J; j++) {
    const int c = 4 * (sl + 32 * j);
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
  if (sl == 0) { atomicAdd(&s_db2[0], ab0); atomicAdd(&s_db2[1], ab1); }
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
  const int G = frm_lanes_per_row(C);
  int gx = cdiv(HW, 8 * 8 * (32 / G));  // >= 8 row groups per warp before the reduction flush
  if (gx < 1) gx = 1;
  if (gx > 296) gx = 296;
  dim3 grid(gx, B);
#define FRB(Jv)                                                                                                              \
  frm_rectify_bwd_kernel<Jv><<<grid, 256, 0, (cudaStream_t)stream>>>(dr1, lddr1, dr2, lddr2, (const bf16*)a, lda, (const bf16*)t, ldt, \
                                                                     w2, cw, sw, da, ldda, (bf16*)dt, lddt, dcw, dw2, db2, HW, C, G)
  const int nj = (C + 127) / 128;
  if (nj <= 1) FRB(1);
  else if (nj == 2) FRB(2);
  else if (nj == 3) FRB(3);
  else FRB(4);
#undef FRB
  LAUNCH_DONE("frm_rectify_bwd");
}
