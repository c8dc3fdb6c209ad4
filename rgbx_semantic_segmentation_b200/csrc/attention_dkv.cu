// EXPERIMENTAL (kernel-level parity green on B200: tests/test_ops_gpu.py::test_attention_dkv_recompute, 9 shapes; not on the default
// path yet - the engine only calls it under CMX_ATTN_DKV_RECOMPUTE=1 until the model-level tests and the bench A/B have run):
// key-major attention backward for dK / dV with the probabilities RECOMPUTED from Q, K and the forward's row
// normaliser instead of being read back from HBM (dual_segformer.py:127-134, backward of softmax(scale Q K^T) V).
//
//   per CTA: one (sample, head), one block of 128 keys, one contiguous range of 128-query tiles
//     S^T [128 keys x 128 queries] = K_blk Q_i^T          tcgen05.mma, TMEM columns   0..127
//     dP^T[128 keys x 128 queries] = V_blk dO_i^T         tcgen05.mma, TMEM columns 128..255
//     P^T  = exp2(scale*log2e * S^T - log2e * lse[q])     (rows >= Nkv and queries >= N give 0)
//     dS^T = scale * P^T .* (dP^T - delta[q])             delta[q] = rowsum(dO .* O)  (cmx_attn_delta)
//     dV  += P^T  dO_i   (A = P^T  from shared memory, B = the dO tile viewed MN-major)   TMEM columns 256..319
//     dK  += dS^T Q_i    (A = dS^T from shared memory, B = the Q  tile viewed MN-major)   TMEM columns 320..383
//   the [128 x 64] dV / dK accumulators stay in tensor memory over all query tiles of the CTA and are added to the
//   fp32 [B*Nkv, 2C] gradient of the kv projection output at the end (fp32 RED: the query range of one (sample, head,
//   key block) is split over several CTAs so that the high-resolution stages fill the GPU).
//
// It replaces the two batched split-K GEMMs dV = P^T dO and dK = dS^T Q (which re-read the stored P and dS,
// 2 x 1.35 GB per training step at MiT-B2 480x640 batch 8).  Warp roles as in attention.cu: warp 0 TMA producer,
// warp 1 MMA issuer, warps 2-9 (two threads per key row, 64 query columns each) exp2 / dS / bf16 staging.
#include "tc_common.cuh"
// one MUFU.EX2 (exp2f() adds a 3-instruction range fix-up; the arguments here are <= ~0 and tiny results may flush to 0)
__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
#include "../../include/cmx_b200.h"
#include <atomic>
#include <string.h>
extern std::atomic<long long> g_cmx_launches;

constexpr int DK_BK = 128;   // keys per CTA
constexpr int DK_BQ = 128;   // queries per tile
constexpr int DK_D = 64;     // head dim
constexpr int DK_NP = 4;                       // column parts of a 128-wide S / dP tile, one group of four warps each
constexpr int DK_SW = DK_NP * 128;              // exp2 / dS threads
constexpr int DK_CPP = 128 / DK_NP;             // columns per part (a multiple of 32)
constexpr int DK_THREADS = 64 + DK_SW;
constexpr uint32_t DK_K_OFF = 0, DK_V_OFF = 16384, DK_Q_OFF = 32768, DK_DO_OFF = 65536, DK_P_OFF = 98304, DK_DS_OFF = 131072,
                   DK_LD_OFF = 163840, DK_BAR_OFF = DK_LD_OFF + 2048;
constexpr uint32_t DK_SMEM = DK_BAR_OFF + 256 + 1024;  // + alignment slack
constexpr uint32_t DK_COL_S = 0, DK_COL_DP = 128, DK_COL_DV = 256, DK_COL_DK = 320;

struct DkvArgs {
  const float* lse;    // [B*heads, N] natural-log row normaliser of the forward
  const float* delta;  // [B*heads, N] rowsum(dO .* O)
  float* dkv;          // [B*Nk, lddkv] fp32, dK at columns h*64.., dV at heads*64 + h*64..
  long lddkv;
  int B, N, Nk, heads;
  int nkb;             // key blocks per (sample, head)
  int qsplit;          // CTAs per (sample, head, key block)
  int tiles_per_split; // query tiles per CTA
  int tiles;           // query tiles per (sample, head)
  float scale_log2e, scale;
};

__global__ void __launch_bounds__(DK_THREADS, 1) attn_dkv_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmDO,
                                                                 const __grid_constant__ CUtensorMap tmKV, DkvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = sb + DK_K_OFF, sV = sb + DK_V_OFF, sQ = sb + DK_Q_OFF, sDO = sb + DK_DO_OFF, sP = sb + DK_P_OFF,
                 sDS = sb + DK_DS_OFF, sLD = sb + DK_LD_OFF;
  const uint32_t bar = sb + DK_BAR_OFF;
  const uint32_t kv_full = bar, qd_full = bar + 8 /*[2]*/, qd_empty = bar + 24 /*[2]*/, s_full = bar + 40, p_full = bar + 48,
                 pds_empty = bar + 56, acc_full = bar + 64, tmem_slot = bar + 72;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // (sample, head), key block and query-tile range of this CTA
  const int split = (int)(blockIdx.x % (unsigned)a.qsplit);
  const int unit = (int)(blockIdx.x / (unsigned)a.qsplit);
  const int jb = unit % a.nkb;
  const int bh = unit / a.nkb;
  const int b = bh / a.heads, h = bh % a.heads;
  const int t_begin = split * a.tiles_per_split;
  int t_end = t_begin + a.tiles_per_split;
  if (t_end > a.tiles) t_end = a.tiles;
  const int ntiles = t_end > t_begin ? t_end - t_begin : 0;
  const int C = a.heads * DK_D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDO)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; i++) { mbar_init(qd_full + 8 * i, 1); mbar_init(qd_empty + 8 * i, 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 1);
    mbar_init(pds_empty, 1);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(kv_full, 2 * DK_BK * 128);
      tma_load_3d(sK, &tmKV, kv_full, h * DK_D, jb * DK_BK, b);        // rows >= Nkv are zero-filled
      tma_load_3d(sV, &tmKV, kv_full, C + h * DK_D, jb * DK_BK, b);
      for (int i = 0; i < ntiles; i++) {
        const int s = i & 1;
        const int q0 = (t_begin + i) * DK_BQ;
        mbar_wait(qd_empty + 8 * s, (((uint32_t)i >> 1) & 1u) ^ 1u);
        mbar_expect_tx(qd_full + 8 * s, 2 * DK_BQ * 128);
        tma_load_3d(sQ + s * 16384, &tmQ, qd_full + 8 * s, h * DK_D, q0, b);
        tma_load_3d(sDO + s * 16384, &tmDO, qd_full + 8 * s, h * DK_D, q0, b);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0 && ntiles > 0) {
      // D fp32, A/B bf16, K-major A; B K-major for the score products, MN-major (bit 16) for the gradient products
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DK_BQ >> 3) << 17) | ((uint32_t)(DK_BK >> 4) << 24);
      constexpr uint32_t idesc_g = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DK_D >> 3) << 17) |
                                   ((uint32_t)(DK_BK >> 4) << 24);
      auto issue_scores = [&](int i) {  // S^T = K Q_i^T, dP^T = V dO_i^T
        const uint32_t q = sQ + (i & 1) * 16384, g = sDO + (i & 1) * 16384;
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
          tc_mma_bf16(tmem + DK_COL_S, umma_desc(sK + ks * 32, 16, 1024), umma_desc(q + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
          tc_mma_bf16(tmem + DK_COL_DP, umma_desc(sV + ks * 32, 16, 1024), umma_desc(g + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(s_full);
      };
      auto issue_grads = [&](int i) {   // dV += P^T dO_i, dK += dS^T Q_i (contraction over the 128 queries of tile i)
        const uint32_t q = sQ + (i & 1) * 16384, g = sDO + (i & 1) * 16384;
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + DK_COL_DV, umma_desc(sP + kb * 16384 + ks * 32, 16, 1024),
                        umma_desc(g + kb * 8192 + ks * 2048, 8192, 1024), idesc_g, (i > 0 || kb > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + DK_COL_DK, umma_desc(sDS + kb * 16384 + ks * 32, 16, 1024),
                        umma_desc(q + kb * 8192 + ks * 2048, 8192, 1024), idesc_g, (i > 0 || kb > 0 || ks > 0) ? 1u : 0u);
      };
      mbar_wait(kv_full, 0);
      for (int i = 0; i < ntiles; i++) {
        mbar_wait(qd_full + 8 * (i & 1), ((uint32_t)i >> 1) & 1u);
        if (i > 0) mbar_wait(p_full, (uint32_t)(i - 1) & 1u);   // S / dP of tile i-1 consumed, P^T / dS^T of tile i-1 staged
        tc_fence_after();
        issue_scores(i);                                        // first, so that the exp2 warps restart as early as possible
        if (i > 0) {
          issue_grads(i - 1);
          tc_commit(pds_empty);                                 // P^T / dS^T staging tiles free again
          tc_commit(qd_empty + 8 * ((i - 1) & 1));              // Q / dO ring slot free again
        }
      }
      mbar_wait(p_full, (uint32_t)(ntiles - 1) & 1u);
      tc_fence_after();
      issue_grads(ntiles - 1);
      tc_commit(acc_full);
    }
  } else if (ntiles > 0) {
    // ============================ exp2 / dS warps (2..9) ============================
    const int qd = warp & 3;               // TMEM lane quarter
    const int part = (warp - 2) >> 2;      // query-column part: [part * DK_CPP, (part + 1) * DK_CPP)
    const int r = qd * 32 + lane;          // key row inside the block
    const int ct = threadIdx.x - 64;       // 0..DK_SW-1
    const uint32_t t_row = tmem + ((uint32_t)(qd * 32) << 16);
    const bool key_ok = jb * DK_BK + r < a.Nk;
    const float sl2 = a.scale_log2e, sc = a.scale;
    const float* lse = a.lse + (long)bh * a.N;
    const float* dlt = a.delta + (long)bh * a.N;
    for (int i = 0; i < ntiles; i++) {
      const int q0 = (t_begin + i) * DK_BQ;
      const uint32_t buf = sLD + (uint32_t)(i & 1) * 1024u;   // double-buffered [128 x lse*log2e][128 x delta] (2 x 512 B)
      // stage the per-query constants of this tile: +inf normaliser for queries >= N makes their probability exactly 0
      if (ct < 256) {
        const int qi = q0 + (ct & 127);
        float val;
        if (ct < 128) val = qi < a.N ? lse[qi] * 1.4426950408889634f : INFINITY;
        else val = qi < a.N ? dlt[qi] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(buf + (ct < 128 ? 0u : 512u) + 4u * (uint32_t)(ct & 127)), "f"(val) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(DK_SW) : "memory");
      mbar_wait(s_full, (uint32_t)i & 1u);                     // S^T and dP^T of tile i are in tensor memory
      if (i > 0) mbar_wait(pds_empty, (uint32_t)(i - 1) & 1u); // the gradient MMAs of tile i-1 have read the staging tiles
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < DK_CPP / 32; c++) {
        const int col0 = part * DK_CPP + c * 32;
        uint32_t sv[32], dv[32];
        tmem_ld32(t_row + DK_COL_S + (uint32_t)col0, sv);
        tmem_ld32(t_row + DK_COL_DP + (uint32_t)col0, dv);
        tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; g++) {
          float l2[8], dl[8];
#pragma unroll
          for (int v4 = 0; v4 < 2; v4++) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l2[4 * v4]), "=f"(l2[4 * v4 + 1]), "=f"(l2[4 * v4 + 2]),
                         "=f"(l2[4 * v4 + 3]) : "r"(buf + 4u * (uint32_t)(col0 + g * 8 + 4 * v4)));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dl[4 * v4]), "=f"(dl[4 * v4 + 1]), "=f"(dl[4 * v4 + 2]),
                         "=f"(dl[4 * v4 + 3]) : "r"(buf + 512u + 4u * (uint32_t)(col0 + g * 8 + 4 * v4)));
          }
          uint32_t pk[4], dk[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int e = g * 8 + 2 * j;
            const float p0 = key_ok ? ex2_approx_f(fmaf(__uint_as_float(sv[e]), sl2, -l2[2 * j])) : 0.f;
            const float p1 = key_ok ? ex2_approx_f(fmaf(__uint_as_float(sv[e + 1]), sl2, -l2[2 * j + 1])) : 0.f;
            const float s0 = sc * p0 * (__uint_as_float(dv[e]) - dl[2 * j]);
            const float s1 = sc * p1 * (__uint_as_float(dv[e + 1]) - dl[2 * j + 1]);
            __nv_bfloat162 hp = __floats2bfloat162_rn(p0, p1), hs = __floats2bfloat162_rn(s0, s1);
            pk[j] = *reinterpret_cast<uint32_t*>(&hp);
            dk[j] = *reinterpret_cast<uint32_t*>(&hs);
          }
          // K-major SWIZZLE_128B staging: 64 queries (128 B) per row and k-block, 16-byte chunk index XOR (row & 7)
          const uint32_t ch = (uint32_t)(((col0 & 63) >> 3) + g);
          const uint32_t off = (uint32_t)(col0 >> 6) * 16384u + (uint32_t)r * 128u + ((ch ^ ((uint32_t)r & 7u)) << 4);
          st_shared_v4(sP + off, pk[0], pk[1], pk[2], pk[3]);
          st_shared_v4(sDS + off, dk[0], dk[1], dk[2], dk[3]);
        }
      }
      tc_fence_before();
      fence_async_smem();
      asm volatile("bar.sync 2, %0;" ::"n"(DK_SW) : "memory");   // staging tiles complete and fenced for the async proxy; TMEM reads done
      if (ct == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: dV / dK accumulators (TMEM) -> fp32 adds into the kv-projection gradient
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int key = jb * DK_BK + r;
    // 2 accumulators x 2 column halves = 4 (accumulator, half) pieces of 32 columns, dealt round-robin to the DK_NP parts
#pragma unroll 1
    for (int piece = part; piece < 4; piece += DK_NP) {   // w = 0: dV (columns C + ...), w = 1: dK
      const int w = piece >> 1, half = piece & 1;
      float* row = a.dkv + ((long)b * a.Nk + key) * a.lddkv + h * DK_D + half * 32;
      uint32_t v[32];
      tmem_ld32(t_row + (w == 0 ? DK_COL_DV : DK_COL_DK) + (uint32_t)(half * 32), v);
      tmem_wait_ld();
      if (key_ok) {
        float* dst = row + (w == 0 ? C : 0);
#pragma unroll
        for (int j = 0; j < 32; j++) atomicAdd(dst + j, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL companion: query-major dQ with recomputed probabilities (no stored P, no dS round trip).
//   persistent CTAs over 128-query tiles ordered by (sample, head); K and V of the (sample, head) stay in shared memory as
//   up to three 128-key blocks (Nkv <= 384); per tile and key block:
//     S  [128 q x 128 k] = Q_i K_blk^T, dP = dO_i V_blk^T            TMEM columns 0..127 / 128..255
//     dS = scale * exp2(scale*log2e*S - log2e*lse[q]) .* (dP - delta[q])   (keys >= Nkv and queries >= N give 0) -> bf16 staging
//     dQ += dS K_blk   (B = the K block viewed MN-major)             TMEM columns 256..319, stored as bf16 per tile
// ------------------------------------------------------------------------------------------------------------------
constexpr int DQ_MAXKB = 3;
constexpr uint32_t DQ_K_OFF = 0, DQ_V_OFF = 49152, DQ_Q_OFF = 98304, DQ_DO_OFF = 131072, DQ_DS_OFF = 163840, DQ_BAR_OFF = 196608;
constexpr uint32_t DQ_SMEM = DQ_BAR_OFF + 256 + 1024;
constexpr uint32_t DQ_COL_S = 0, DQ_COL_DP = 128, DQ_COL_DQ = 256;

struct DqArgs {
  const float* lse;
  const float* delta;
  bf16* dq;
  long lddq;
  int B, N, Nk, heads;
  int nkb;
  int tiles_per_bh;
  long total_tiles;
  float scale_log2e, scale;
};

__global__ void __launch_bounds__(DK_THREADS, 1) attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmDO,
                                                                const __grid_constant__ CUtensorMap tmKV, DqArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = sb + DQ_K_OFF, sV = sb + DQ_V_OFF, sQ = sb + DQ_Q_OFF, sDO = sb + DQ_DO_OFF, sDS = sb + DQ_DS_OFF;
  const uint32_t bar = sb + DQ_BAR_OFF;
  const uint32_t kv_full = bar, kv_empty = bar + 8, qd_full = bar + 16 /*[2]*/, qd_empty = bar + 32 /*[2]*/, s_full = bar + 48,
                 p_full = bar + 56, ds_empty = bar + 64, o_full = bar + 72, o_empty = bar + 80, tmem_slot = bar + 88;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long per = (a.total_tiles + gridDim.x - 1) / gridDim.x;
  const long t_begin = (long)blockIdx.x * per;
  long t_end = t_begin + per;
  if (t_end > a.total_tiles) t_end = a.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;
  const int nkb = a.nkb;
  const int C = a.heads * DK_D;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDO)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; i++) { mbar_init(qd_full + 8 * i, 1); mbar_init(qd_empty + 8 * i, 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 1);
    mbar_init(ds_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, 4 * DK_NP);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int group = -1;
      long cur_bh = -1;
      for (int i = 0; i < ntiles; i++) {
        const long t = t_begin + i;
        const long bh = t / a.tiles_per_bh;
        const int q0 = (int)(t % a.tiles_per_bh) * DK_BQ;
        const int b = (int)(bh / a.heads), h = (int)(bh % a.heads);
        if (bh != cur_bh) {
          cur_bh = bh;
          group++;
          if (group > 0) mbar_wait(kv_empty, (uint32_t)(group - 1) & 1u);   // every MMA that read the old K / V retired
          mbar_expect_tx(kv_full, (uint32_t)(2 * nkb) * 16384u);
          for (int kb = 0; kb < nkb; kb++) {
            tma_load_3d(sK + kb * 16384, &tmKV, kv_full, h * DK_D, kb * DK_BK, b);
            tma_load_3d(sV + kb * 16384, &tmKV, kv_full, C + h * DK_D, kb * DK_BK, b);
          }
        }
        const int s = i & 1;
        mbar_wait(qd_empty + 8 * s, (((uint32_t)i >> 1) & 1u) ^ 1u);
        mbar_expect_tx(qd_full + 8 * s, 2 * DK_BQ * 128);
        tma_load_3d(sQ + s * 16384, &tmQ, qd_full + 8 * s, h * DK_D, q0, b);
        tma_load_3d(sDO + s * 16384, &tmDO, qd_full + 8 * s, h * DK_D, q0, b);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DK_BK >> 3) << 17) | ((uint32_t)(DK_BQ >> 4) << 24);
      constexpr uint32_t idesc_g = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(DK_D >> 3) << 17) |
                                   ((uint32_t)(DK_BQ >> 4) << 24);
      auto tile_bh = [&](int i) { return (t_begin + i) / a.tiles_per_bh; };
      // dQ_i += dS(i, blk) K_blk, plus the barriers that its retirement releases
      auto issue_dq = [&](int i, int blk) {
        if (blk == 0 && i > 0) mbar_wait(o_empty, (uint32_t)(i - 1) & 1u);   // epilogue of tile i-1 drained the accumulator
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; kb++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + DQ_COL_DQ, umma_desc(sDS + kb * 16384 + ks * 32, 16, 1024),
                        umma_desc(sK + blk * 16384 + kb * 8192 + ks * 2048, 8192, 1024), idesc_g, (blk > 0 || kb > 0 || ks > 0) ? 1u : 0u);
        tc_commit(ds_empty);
        if (blk == nkb - 1) {
          tc_commit(o_full);
          tc_commit(qd_empty + 8 * (i & 1));
          if (i + 1 == ntiles || tile_bh(i + 1) != tile_bh(i)) tc_commit(kv_empty);
        }
      };
      int st = 0, group = 0, pi = 0, pblk = 0;
      mbar_wait(kv_full, 0);
      for (int i = 0; i < ntiles; i++) {
        for (int blk = 0; blk < nkb; blk++) {
          if (st > 0) {
            mbar_wait(p_full, (uint32_t)(st - 1) & 1u);   // S / dP of the previous step consumed, its dS staged
            issue_dq(pi, pblk);
          }
          if (blk == 0) {
            if (i > 0 && tile_bh(i) != tile_bh(i - 1)) {
              group++;
              mbar_wait(kv_full, (uint32_t)group & 1u);
            }
            mbar_wait(qd_full + 8 * (i & 1), ((uint32_t)i >> 1) & 1u);
          }
          tc_fence_after();
          const uint32_t q = sQ + (i & 1) * 16384, g = sDO + (i & 1) * 16384;
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + DQ_COL_S, umma_desc(q + ks * 32, 16, 1024), umma_desc(sK + blk * 16384 + ks * 32, 16, 1024), idesc_s,
                        ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + DQ_COL_DP, umma_desc(g + ks * 32, 16, 1024), umma_desc(sV + blk * 16384 + ks * 32, 16, 1024), idesc_s,
                        ks > 0 ? 1u : 0u);
          tc_commit(s_full);
          pi = i;
          pblk = blk;
          st++;
        }
      }
      mbar_wait(p_full, (uint32_t)(st - 1) & 1u);
      issue_dq(pi, pblk);
    }
  } else {
    // ============================ exp2 / dS warps (2..9) + dQ epilogue ============================
    const int qd = warp & 3;
    const int part = (warp - 2) >> 2;      // key-column part of the block: [part * DK_CPP, (part + 1) * DK_CPP)
    const int r = qd * 32 + lane;          // query row inside the tile
    const int ct = threadIdx.x - 64;
    const uint32_t t_row = tmem + ((uint32_t)(qd * 32) << 16);
    const float sl2 = a.scale_log2e, sc = a.scale;
    int st = 0;
    for (int i = 0; i < ntiles; i++) {
      const long t = t_begin + i;
      const long bh = t / a.tiles_per_bh;
      const int q0 = (int)(t % a.tiles_per_bh) * DK_BQ;
      const int b = (int)(bh / a.heads), h = (int)(bh % a.heads);
      const bool q_ok = q0 + r < a.N;
      const float l2 = q_ok ? a.lse[bh * a.N + q0 + r] * 1.4426950408889634f : INFINITY;
      const float dl = q_ok ? a.delta[bh * a.N + q0 + r] : 0.f;
      for (int blk = 0; blk < nkb; blk++, st++) {
        mbar_wait(s_full, (uint32_t)st & 1u);
        if (st > 0) mbar_wait(ds_empty, (uint32_t)(st - 1) & 1u);   // the dQ MMAs of the previous step have read the staging tile
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < DK_CPP / 32; c++) {
          const int col0 = part * DK_CPP + c * 32;
          const int key0 = blk * DK_BK + col0;
          uint32_t sv[32], dv[32];
          tmem_ld32(t_row + DQ_COL_S + (uint32_t)col0, sv);
          tmem_ld32(t_row + DQ_COL_DP + (uint32_t)col0, dv);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint32_t dk[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int e = g * 8 + 2 * j;
              const float p0 = key0 + e < a.Nk ? ex2_approx_f(fmaf(__uint_as_float(sv[e]), sl2, -l2)) : 0.f;
              const float p1 = key0 + e + 1 < a.Nk ? ex2_approx_f(fmaf(__uint_as_float(sv[e + 1]), sl2, -l2)) : 0.f;
              __nv_bfloat162 hs = __floats2bfloat162_rn(sc * p0 * (__uint_as_float(dv[e]) - dl), sc * p1 * (__uint_as_float(dv[e + 1]) - dl));
              dk[j] = *reinterpret_cast<uint32_t*>(&hs);
            }
            const uint32_t ch = (uint32_t)(((col0 & 63) >> 3) + g);
            st_shared_v4(sDS + (uint32_t)(col0 >> 6) * 16384u + (uint32_t)r * 128u + ((ch ^ ((uint32_t)r & 7u)) << 4), dk[0], dk[1], dk[2], dk[3]);
          }
        }
        tc_fence_before();
        fence_async_smem();
        asm volatile("bar.sync 1, %0;" ::"n"(DK_SW) : "memory");
        if (ct == 0) mbar_arrive(p_full);
      }
      // ---- epilogue: dQ tile TMEM -> bf16 -> global; each thread of the pair takes 32 of the 64 columns
      mbar_wait(o_full, (uint32_t)i & 1u);
      tc_fence_after();
      {
        constexpr int OC = DK_D / DK_NP;   // dQ columns per part
        uint32_t v[OC];
        tmem_ld_n<OC>(t_row + DQ_COL_DQ + (uint32_t)(part * OC), v);
        tmem_wait_ld();
        if (q_ok) {
          bf16* dst = a.dq + ((long)b * a.N + q0 + r) * a.lddq + h * DK_D + part * OC;
#pragma unroll
          for (int g = 0; g < OC / 8; g++) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = __uint_as_float(v[g * 8 + j]);
            store8(dst + g * 8, f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// delta[(b*heads + h)*N + n] = sum_j dO[b,n,h*64+j] * O[b,n,h*64+j]   (one warp per token, two elements per lane and head)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ d_o, long lddo, const bf16* __restrict__ o, long ldo,
                                                         float* __restrict__ delta, long rows, int N, int heads) {
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long b = row / N, n = row % N;
  const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(d_o + row * lddo);
  const __nv_bfloat162* c = reinterpret_cast<const __nv_bfloat162*>(o + row * ldo);
  for (int h = 0; h < heads; h++) {
    const float2 x = __bfloat1622float2(a[h * 32 + lane]), y = __bfloat1622float2(c[h * 32 + lane]);
    float s = fmaf(x.x, y.x, x.y * y.y);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) delta[(b * heads + h) * N + n] = s;
  }
}

CMX_API int cmx_attn_delta(const void* d_o, int64_t lddo, const void* o, int64_t ldo, float* delta, int B, int N, int heads,
                           void* stream) {
  CMX_REQUIRE(d_o && o && delta, "attn_delta: null operand");
  CMX_REQUIRE(lddo % 2 == 0 && ldo % 2 == 0 && lddo >= heads * DK_D && ldo >= heads * DK_D, "attn_delta: head_dim must be 64");
  CMX_REQUIRE(((uintptr_t)d_o & 3) == 0 && ((uintptr_t)o & 3) == 0, "attn_delta: pointers must be 4-byte aligned");
  const long rows = (long)B * N;
  if (rows == 0) return 0;
  attn_delta_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)d_o, lddo, (const bf16*)o, ldo, delta, rows, N,
                                                                               heads);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_delta_kernel");
  return 0;
}

CMX_API int cmx_attn_dkv(const void* q, int64_t ldq, const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const float* lse,
                         const float* delta, float* dkv_acc, int64_t lddkv, int B, int N, int Nk, int heads, float scale,
                         void* stream) {
  CMX_REQUIRE(q && d_o && kv && lse && delta && dkv_acc, "attn_dkv: null operand");
  CMX_REQUIRE(Nk >= 1, "attn_dkv: Nkv=%d", Nk);
  CMX_REQUIRE(ldq % 8 == 0 && lddo % 8 == 0 && ldkv % 8 == 0, "attn_dkv: leading dims must be multiples of 8");
  CMX_REQUIRE(ldq >= heads * DK_D && lddo >= heads * DK_D && ldkv >= 2 * heads * DK_D && lddkv >= 2 * heads * DK_D,
              "attn_dkv: head_dim must be 64");
  CMX_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)d_o & 15) == 0 && ((uintptr_t)kv & 15) == 0, "attn_dkv: pointers must be 16-byte aligned");
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmDO, tmKV;
  int rc = cmx_make_map3(&tmQ, q, (uint64_t)heads * DK_D, (uint64_t)N, (uint64_t)B, (uint64_t)ldq, (uint64_t)N * ldq, DK_D, DK_BQ);
  if (rc) return rc;
  rc = cmx_make_map3(&tmDO, d_o, (uint64_t)heads * DK_D, (uint64_t)N, (uint64_t)B, (uint64_t)lddo, (uint64_t)N * lddo, DK_D, DK_BQ);
  if (rc) return rc;
  rc = cmx_make_map3(&tmKV, kv, (uint64_t)2 * heads * DK_D, (uint64_t)Nk, (uint64_t)B, (uint64_t)ldkv, (uint64_t)Nk * ldkv, DK_D, DK_BK);
  if (rc) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  DkvArgs a;
  a.lse = lse; a.delta = delta; a.dkv = dkv_acc; a.lddkv = lddkv;
  a.B = B; a.N = N; a.Nk = Nk; a.heads = heads;
  a.nkb = (int)cdiv(Nk, DK_BK);
  a.tiles = (int)cdiv(N, DK_BQ);
  const long units = (long)B * heads * a.nkb;
  long qs = units >= sms ? 1 : sms / units;        // fill the GPU once; every split keeps >= 1 query tile
  if (qs > a.tiles) qs = a.tiles;
  a.tiles_per_split = (int)cdiv(a.tiles, qs);
  a.qsplit = (int)cdiv(a.tiles, a.tiles_per_split);
  a.scale_log2e = scale * 1.4426950408889634f;
  a.scale = scale;
  const long grid = units * a.qsplit;
  CMX_REQUIRE(grid < (1l << 31), "attn_dkv: grid too large");
  static thread_local PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DK_SMEM);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute(attn_dkv): %s", cudaGetErrorString(e));
    attr_once.mark();
  }
  attn_dkv_kernel<<<(unsigned)grid, DK_THREADS, DK_SMEM, st>>>(tmQ, tmDO, tmKV, a);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_dkv_kernel");
  return 0;
}

CMX_API int cmx_attn_dq(const void* q, int64_t ldq, const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const float* lse,
                        const float* delta, void* dq, int64_t lddq, int B, int N, int Nk, int heads, float scale,
                        int64_t kv_sample_rows, void* stream) {
  CMX_REQUIRE(q && d_o && kv && lse && delta && dq, "attn_dq: null operand");
  if (kv_sample_rows <= 0) kv_sample_rows = Nk;   // key rows per sample in kv (> Nk when this call covers one key chunk only)
  CMX_REQUIRE(Nk >= 1 && Nk <= DQ_MAXKB * DK_BK, "attn_dq: Nkv=%d unsupported (max %d) - use the unfused path", Nk, DQ_MAXKB * DK_BK);
  CMX_REQUIRE(ldq % 8 == 0 && lddo % 8 == 0 && ldkv % 8 == 0 && lddq % 8 == 0, "attn_dq: leading dims must be multiples of 8");
  CMX_REQUIRE(ldq >= heads * DK_D && lddo >= heads * DK_D && ldkv >= 2 * heads * DK_D && lddq >= heads * DK_D, "attn_dq: head_dim must be 64");
  CMX_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)d_o & 15) == 0 && ((uintptr_t)kv & 15) == 0 && ((uintptr_t)dq & 15) == 0,
              "attn_dq: pointers must be 16-byte aligned");
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmDO, tmKV;
  int rc = cmx_make_map3(&tmQ, q, (uint64_t)heads * DK_D, (uint64_t)N, (uint64_t)B, (uint64_t)ldq, (uint64_t)N * ldq, DK_D, DK_BQ);
  if (rc) return rc;
  rc = cmx_make_map3(&tmDO, d_o, (uint64_t)heads * DK_D, (uint64_t)N, (uint64_t)B, (uint64_t)lddo, (uint64_t)N * lddo, DK_D, DK_BQ);
  if (rc) return rc;
  rc = cmx_make_map3(&tmKV, kv, (uint64_t)2 * heads * DK_D, (uint64_t)Nk, (uint64_t)B, (uint64_t)ldkv, (uint64_t)kv_sample_rows * ldkv, DK_D, DK_BK);
  if (rc) return rc;
  DqArgs a;
  a.lse = lse; a.delta = delta; a.dq = (bf16*)dq; a.lddq = lddq;
  a.B = B; a.N = N; a.Nk = Nk; a.heads = heads;
  a.nkb = (int)cdiv(Nk, DK_BK);
  a.tiles_per_bh = (int)cdiv(N, DK_BQ);
  a.total_tiles = (long)a.tiles_per_bh * B * heads;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.scale = scale;
  static thread_local PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQ_SMEM);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute(attn_dq): %s", cudaGetErrorString(e));
    attr_once.mark();
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long grid = a.total_tiles < sms ? a.total_tiles : sms;
  attn_dq_kernel<<<(unsigned)grid, DK_THREADS, DQ_SMEM, st>>>(tmQ, tmDO, tmKV, a);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_dq_kernel");
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Key-chunked attention (Nkv above the resident-K/V limit of the fused kernels, e.g. 880 / 920 at 720x1280): the fused
// forward runs once per chunk of <= 320 keys and writes that chunk's normalised output and log-sum-exp; the softmax over all
// keys is then   O = sum_c exp(lse_c - lse) O_c,   lse = log sum_c exp(lse_c)   (exact; no [N, Nkv] tensor ever exists).
// thread = 8 channels of one (row, head); parts are stacked [nc][rows, ld] / [nc][B*heads*N]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_combine_kernel(const bf16* __restrict__ o_parts, long part_stride, long ldp,
                                                          const float* __restrict__ lse_parts, long lse_stride, int nc,
                                                          bf16* __restrict__ o, long ldo, float* __restrict__ lse, int B, int N, int heads) {
  pdl_trigger();
  const long total = (long)B * N * heads * 8;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = (int)(idx & 7);
  const int h = (int)((idx >> 3) % heads);
  const long row = idx / ((long)heads * 8);
  const long b = row / N, n = row % N;
  const long li = (b * heads + h) * N + n;
  float m = -INFINITY;
  for (int c = 0; c < nc; c++) m = fmaxf(m, lse_parts[c * lse_stride + li]);
  float den = 0.f;
  for (int c = 0; c < nc; c++) den += expf(lse_parts[c * lse_stride + li] - m);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) acc[i] = 0.f;
  for (int c = 0; c < nc; c++) {
    const float w = expf(lse_parts[c * lse_stride + li] - m) / den;
    float t[8];
    load8(o_parts + c * part_stride + row * ldp + h * 64 + v * 8, t);
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] += w * t[i];
  }
  store8(o + row * ldo + h * 64 + v * 8, acc);
  if (v == 0) lse[li] = m + logf(den);
}
CMX_API int cmx_attn_combine(const void* o_parts, int64_t part_stride, int64_t ldp, const float* lse_parts, int64_t lse_stride, int nc,
                             void* o, int64_t ldo, float* lse, int B, int N, int heads, void* stream) {
  CMX_REQUIRE(o_parts && lse_parts && o && lse && nc >= 1, "attn_combine: null operand");
  CMX_REQUIRE(ldp % 8 == 0 && ldo % 8 == 0 && part_stride % 8 == 0, "attn_combine: strides must be multiples of 8");
  const long total = (long)B * N * heads * 8;
  if (total == 0) return 0;
  attn_combine_kernel<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)o_parts, part_stride, ldp, lse_parts,
                                                                                  lse_stride, nc, (bf16*)o, ldo, lse, B, N, heads);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_combine");
  return 0;
}

// out = sum of nc stacked bf16 tensors (fp32 accumulation): the per-chunk dQ partials of the key-chunked backward
__global__ void __launch_bounds__(256) sum_parts_kernel(const bf16* __restrict__ parts, long part_stride, int nc, bf16* __restrict__ out, long n8) {
  pdl_trigger();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = 0.f;
  for (int c = 0; c < nc; c++) {
    float t[8];
    load8(parts + c * part_stride + i * 8, t);
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] += t[k];
  }
  store8(out + i * 8, acc);
}
CMX_API int cmx_sum_parts_bf16(const void* parts, int64_t part_stride, int nc, void* out, int64_t n, void* stream) {
  CMX_REQUIRE(parts && out && nc >= 1 && n % 8 == 0 && part_stride % 8 == 0, "sum_parts: n and the part stride must be multiples of 8");
  if (n == 0) return 0;
  sum_parts_kernel<<<(unsigned)cdiv(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)parts, part_stride, nc, (bf16*)out, n / 8);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("sum_parts");
  return 0;
}
