// Fused spatial-reduction self-attention forward (dual_segformer.py:127-134): O = softmax(scale * Q K^T) V per
// (sample, head), head_dim = 64, Nkv <= 320 (Nkv = 300 at every MiT stage for 480x640 inputs), flash style:
// the N x Nkv score matrix lives only in tensor memory / registers.
//
//   * persistent CTAs walk contiguous ranges of 128-row Q tiles ordered by (sample, head): the K and V tiles of a
//     (sample, head) are TMA-loaded once ([320 x 64] bf16 each, SWIZZLE_128B, rows >= Nkv zero-filled) and stay in
//     shared memory; Q tiles stream through a 2-deep TMA ring.
//   * warp 1 issues tcgen05.mma: S[128 x 320] = Q K^T (two N=160 UMMAs x 4 k-steps, fp32 in TMEM columns 0..319) and
//     O[128 x 64] = P V (A = P from shared memory, B = V viewed MN-major - the same bytes as the K-major tile -
//     TMEM columns 320..383).  The QK^T of tile i+1 is issued before the epilogue of tile i finishes.
//   * warps 2-17 (16 softmax warps = 4 TMEM lane quarters x 4 column parts of 80 columns, walked as 16-column chunks with
//     software-pipelined tcgen05.ld): pass 1 row max, pass 2 row sum of exp2 (training) or unnormalised bf16 P + row sum
//     (inference / recompute backward / key chunks: the 64 output columns are scaled instead), pass 3 (training only) recomputes
//     the exponentials with the normaliser folded in and writes the NORMALISED bf16 P once into the K-major SWIZZLE_128B tile
//     that feeds BOTH the P V MMA and the TMA bulk store of the probabilities the backward pass consumes (issued by lane 1 of
//     warp 0, which also releases the tile).  Row reductions stay inside a quarter (128-thread named barriers); "P complete"
//     is one mbarrier arrive per warp.
#include "tc_common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>
#include <string.h>
extern std::atomic<long long> g_cmx_launches;

constexpr int AT_BM = 128;      // query rows per tile
constexpr int AT_D = 64;        // head dim
constexpr int AT_NK = 320;      // padded key count (5 k-blocks of 64)
#ifndef AT_NP
#define AT_NP 4
#endif
// softmax warps per TMEM lane quarter (each takes a contiguous range of the ten 32-column chunks of S).  The softmax warps are
// instruction-issue / latency bound (ncu: 43-49 % issue activity with 8 of them = 2 per scheduler), so 16 warps split the columns
constexpr int AT_SW = 4 * AT_NP;             // softmax warps
constexpr int AT_ST = 32 * AT_SW;            // softmax threads (named-barrier count)
constexpr int AT_THREADS = 64 + AT_ST;
constexpr uint32_t AT_K_OFF = 0, AT_V_OFF = 40960, AT_Q_OFF = 81920, AT_P_OFF = 114688, AT_RED_OFF = 196608;
constexpr uint32_t AT_RED2 = AT_NP * 512u;   // second reduction array (row sums) behind the first (row max / dot)
constexpr uint32_t AT_BAR_OFF = AT_RED_OFF + 2 * AT_RED2;
constexpr uint32_t AT_SMEM = AT_BAR_OFF + 256 + 1024;  // + alignment slack
constexpr uint32_t AT_O_COL = 320;

__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU.EX2 (2 ulp; the result is rounded to bf16 anyway)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnArgs {
  bf16* o;
  long ldo;
  float* lse;
  int B, N, Nk, heads;
  int tiles_per_bh;
  long total_tiles;
  float scale_log2e;
  float scale;
  int store_p;
};

// MODE 0: forward.  MODE 1: backward core - the same pipeline with  A1 = dO, B1 = V  (dP = dO V^T in TMEM),
// P TMA-loaded into the staging tile, dS = scale * P .* (dP - rowsum(P .* dP)) computed in place and TMA-stored
// (it feeds the split-K dK = dS^T Q GEMM), and the second MMA  dQ = dS K  (B2 = K viewed MN-major).
template <int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                             const __grid_constant__ CUtensorMap tmKV,
                                                             const __grid_constant__ CUtensorMap tmP,
                                                             const __grid_constant__ CUtensorMap tmDS, AttnArgs a) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = sb + AT_K_OFF, sV = sb + AT_V_OFF, sQ = sb + AT_Q_OFF, sP = sb + AT_P_OFF, sRed = sb + AT_RED_OFF;
  const uint32_t bar = sb + AT_BAR_OFF;
  const uint32_t kv_full = bar, kv_empty = bar + 8, q_full = bar + 16 /*[2]*/, q_empty = bar + 32 /*[2]*/, s_full = bar + 48,
                 p_full = bar + 56, o_full = bar + 64, o_empty = bar + 72, p_in = bar + 80, tmem_slot = bar + 88, p_free = bar + 96;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous tile range of this CTA
  const long per = (a.total_tiles + gridDim.x - 1) / gridDim.x;
  const long t_begin = (long)blockIdx.x * per;
  long t_end = t_begin + per;
  if (t_end > a.total_tiles) t_end = a.total_tiles;
  const int ntiles = t_end > t_begin ? (int)(t_end - t_begin) : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; i++) { mbar_init(q_full + 8 * i, 1); mbar_init(q_empty + 8 * i, 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, MODE == 0 ? AT_SW : 1);
    mbar_init(p_free, 1);
    mbar_init(o_full, 1);
    mbar_init(o_empty, AT_SW);
    mbar_init(p_in, 1);
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
  pdl_wait();  // PDL: the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int group = -1;
      long cur_bh = -1;
      for (int i = 0; i < ntiles; i++) {
        const long t = t_begin + i;
        const long bh = t / a.tiles_per_bh;
        const int q0 = (int)(t % a.tiles_per_bh) * AT_BM;
        const int b = (int)(bh / a.heads), h = (int)(bh % a.heads);
        if (bh != cur_bh) {
          cur_bh = bh;
          group++;
          if (group > 0) mbar_wait(kv_empty, (uint32_t)(group - 1) & 1u);  // all MMAs that read the old K/V retired
          mbar_expect_tx(kv_full, 2 * AT_NK * 128);
          const int C = a.heads * AT_D;
          // first-MMA operand (K-major view) goes to slot sK, second-MMA operand (MN-major view) to slot sV:
          // forward (K, V); backward (V, K)
          const int c1 = MODE == 0 ? h * AT_D : C + h * AT_D;
          const int c2 = MODE == 0 ? C + h * AT_D : h * AT_D;
          tma_load_3d(sK, &tmKV, kv_full, c1, 0, b);
          tma_load_3d(sK + 160 * 128, &tmKV, kv_full, c1, 160, b);
          tma_load_3d(sV, &tmKV, kv_full, c2, 0, b);
          tma_load_3d(sV + 160 * 128, &tmKV, kv_full, c2, 160, b);
        }
        const int s = i & 1;
        mbar_wait(q_empty + 8 * s, (((uint32_t)i >> 1) & 1u) ^ 1u);
        mbar_expect_tx(q_full + 8 * s, AT_BM * 128);
        tma_load_3d(sQ + s * 16384, &tmQ, q_full + 8 * s, h * AT_D, q0, b);
      }
    } else if (MODE == 0 && lane == 1 && a.store_p) {
      // ---- P store (training forward): once the sixteen softmax warps have completed tile i's normalised probabilities, TMA-store
      // them for the backward pass and release the staging tile (p_free) when the copy engine has finished reading it
      for (int i = 0; i < ntiles; i++) {
        const long t = t_begin + i;
        const long bh = t / a.tiles_per_bh;
        const int q0 = (int)(t % a.tiles_per_bh) * AT_BM;
        mbar_wait(p_full, (uint32_t)i & 1u);
#pragma unroll
        for (int kb = 0; kb < AT_NK / 64; kb++)
          if (kb * 64 < a.Nk)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmP)), "r"(sP + kb * 16384), "r"(kb * 64), "r"(q0), "r"((int)bh)
                         : "memory");
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(p_free);
      }
      tma_store_wait_all();
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(160 >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(AT_D >> 3) << 17) |
                                   ((uint32_t)(AT_BM >> 4) << 24);
      auto issue_qk = [&](int i) {  // S = Q_i K^T
        const uint32_t q = sQ + (i & 1) * 16384;
#pragma unroll
        for (int hf = 0; hf < 2; hf++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + hf * 160, umma_desc(q + ks * 32, 16, 1024), umma_desc(sK + hf * 160 * 128 + ks * 32, 16, 1024),
                        idesc_s, ks > 0 ? 1u : 0u);
        tc_commit(s_full);
        tc_commit(q_empty + 8 * (i & 1));
      };
      int group = 0;
      mbar_wait(kv_full, 0);
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      for (int i = 0; i < ntiles; i++) {
        const long bh = (t_begin + i) / a.tiles_per_bh;
        const bool last_of_group = (i + 1 == ntiles) || ((t_begin + i + 1) / a.tiles_per_bh != bh);
        mbar_wait(p_full, (uint32_t)i & 1u);                        // P_i is in shared memory, S is free again
        if (i > 0) mbar_wait(o_empty, (uint32_t)(i - 1) & 1u);      // epilogue of tile i-1 has drained O
        tc_fence_after();
        const int nkb_mma = MODE == 0 ? AT_NK / 64 : (a.Nk + 63) / 64;  // backward: only the k-blocks that were loaded
#pragma unroll 1
        for (int kb = 0; kb < nkb_mma; kb++)
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_bf16(tmem + AT_O_COL, umma_desc(sP + kb * 16384 + ks * 32, 16, 1024),
                        umma_desc(sV + kb * 8192 + ks * 2048, 8192, 1024), idesc_o, (kb > 0 || ks > 0) ? 1u : 0u);
        tc_commit(o_full);
        if (last_of_group) tc_commit(kv_empty);
        if (i + 1 < ntiles) {
          if (last_of_group) {
            group++;
            mbar_wait(kv_full, (uint32_t)group & 1u);
          }
          mbar_wait(q_full + 8 * ((i + 1) & 1), ((uint32_t)(i + 1) >> 1) & 1u);
          tc_fence_after();
          issue_qk(i + 1);
        }
      }
    }
  } else {
    // ============================ softmax + epilogue (warps 2..17) ============================
    const int q = warp & 3;               // TMEM lane quarter
    const int part = (warp - 2) >> 2;     // column part of this warp: chunks [cb, ce) of the ten 32-column chunks
    const int cb = (10 * part) / AT_NP, ce = (10 * (part + 1)) / AT_NP;
    const int r = q * 32 + lane;          // row inside the tile
    const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16);
    const float sl2 = a.scale_log2e;
    float o_scale = 1.f;
    for (int i = 0; i < ntiles; i++) {
      const long t = t_begin + i;
      const long bh = t / a.tiles_per_bh;
      const int q0 = (int)(t % a.tiles_per_bh) * AT_BM;
      const int b = (int)(bh / a.heads), h = (int)(bh % a.heads);
      if (MODE == 1) {
        const int nkb = (a.Nk + 63) / 64;
        if (threadIdx.x == 64) {
          tma_store_wait_read<0>();                       // dS store of the previous tile has read the staging tile
          mbar_expect_tx(p_in, (uint32_t)nkb * 16384u);
          for (int kb = 0; kb < nkb; kb++) tma_load_3d(sP + kb * 16384, &tmP, p_in, kb * 64, q0, (int)bh);
        }
        mbar_wait(s_full, (uint32_t)i & 1u);               // dP = dO V^T is in tensor memory
        tc_fence_after();
        mbar_wait(p_in, (uint32_t)i & 1u);                 // P tile landed
        const int ncg = ((a.Nk + 7) >> 3);                 // 8-column groups that can hold non-zero probabilities
        // ---- pass 1: D = rowsum(P .* dP) over this thread's 160 columns
        float dot = 0.f;
#pragma unroll 1
        for (int c = cb; c < ce; c++) {
          const int col0 = c * 32;
          if ((col0 >> 3) >= ncg) break;
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)col0, v);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int col = col0 + g * 8;
            const uint32_t kb = (uint32_t)col >> 6, ch = ((uint32_t)col & 63u) >> 3;
            uint32_t w[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                         : "r"(sP + kb * 16384 + (uint32_t)r * 128 + ((ch ^ ((uint32_t)r & 7u)) << 4)));
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 pf = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[j]));
              dot = fmaf(pf.x, __uint_as_float(v[g * 8 + 2 * j]), dot);
              dot = fmaf(pf.y, __uint_as_float(v[g * 8 + 2 * j + 1]), dot);
            }
          }
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sRed + 4u * (part * 128 + r)), "f"(dot) : "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(AT_ST) : "memory");
        float D = 0.f;
#pragma unroll
        for (int pp = 0; pp < AT_NP; pp++) {
          float dpart;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dpart) : "r"(sRed + 4u * (pp * 128 + r)));
          D += dpart;
        }
        // ---- pass 2: dS = scale * P .* (dP - D), written in place of P
#pragma unroll 1
        for (int c = cb; c < ce; c++) {
          const int col0 = c * 32;
          if ((col0 >> 3) >= ncg) break;
          uint32_t v[32];
          tmem_ld32(t_row + (uint32_t)col0, v);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int col = col0 + g * 8;
            const uint32_t kb = (uint32_t)col >> 6, ch = ((uint32_t)col & 63u) >> 3;
            const uint32_t ad = sP + kb * 16384 + (uint32_t)r * 128 + ((ch ^ ((uint32_t)r & 7u)) << 4);
            uint32_t w[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(ad));
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 pf = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[j]));
              __nv_bfloat162 h2 = __floats2bfloat162_rn(a.scale * pf.x * (__uint_as_float(v[g * 8 + 2 * j]) - D),
                                                        a.scale * pf.y * (__uint_as_float(v[g * 8 + 2 * j + 1]) - D));
              w[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            st_shared_v4(ad, w[0], w[1], w[2], w[3]);
          }
        }
        tc_fence_before();
        fence_async_smem();
        asm volatile("bar.sync 3, %0;" ::"n"(AT_ST) : "memory");
        if (threadIdx.x == 64) {
          mbar_arrive(p_full);
          for (int kb = 0; kb < nkb; kb++)
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmDS)), "r"(sP + kb * 16384), "r"(kb * 64), "r"(q0), "r"((int)bh)
                         : "memory");
          tma_store_commit();
        }
      } else {
        // ---- forward softmax.  Every warp owns the 128 rows of its TMEM lane quarter q and 80 of the 320 score columns (part),
        // walked as five 16-column chunks with software-pipelined tcgen05.ld (the load of chunk k+1 is in flight while chunk k
        // is processed: ncu showed the warps stalled on the scoreboard of the TMEM / shared-memory loads, 30 % issue activity).
        // Row reductions go through shared memory between the four warps of a quarter only (128-thread named barriers), and the
        // "P complete" signal is one mbarrier arrive per warp - no CTA-wide barrier is left in the tile loop.
        mbar_wait(s_full, (uint32_t)i & 1u);
        tc_fence_after();
        const int colb = part * 80;
        const uint32_t t_base = t_row + (uint32_t)colb;
        const int bar_max = 1 + q, bar_sum = 5 + q;
        // AT_WALK(BODY): BODY(values[16], first column) for this warp's chunks below Nkv; two register buffers with fixed roles,
        // the loop walks chunk PAIRS and is not unrolled (macros, not lambdas: the buffers must stay in registers)
#define AT_WALK(BODY)                                                                              \
  {                                                                                                \
    uint32_t v0[16], v1[16];                                                                       \
    tmem_ld_n<16>(t_base, v0);                                                                     \
    _Pragma("unroll 1") for (int k = 0; k < 5; k += 2) {                                           \
      const int col0 = colb + 16 * k;                                                              \
      /* the loads are unconditional (a chunk past this warp's range or past Nkv is read and ignored: the columns exist) */ \
      tmem_wait_ld_dep16(v0);                                                                      \
      tmem_ld_n<16>(t_base + (uint32_t)(16 * (k + 1)), v1);                                        \
      if (col0 < a.Nk) { BODY(v0, col0) }                                                          \
      tmem_wait_ld_dep16(v1);                                                                      \
      if (k + 2 < 5) tmem_ld_n<16>(t_base + (uint32_t)(16 * (k + 2)), v0);                         \
      if (k + 1 < 5 && col0 + 16 < a.Nk) { BODY(v1, (col0 + 16)) }                                 \
    }                                                                                              \
  }
#define AT_BODY_MAX(v, c0)                                                                         \
  if ((c0) + 16 <= a.Nk) {                                                                         \
    _Pragma("unroll") for (int j = 0; j < 16; j++) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j])); \
  } else {                                                                                         \
    _Pragma("unroll") for (int j = 0; j < 16; j++)                                                 \
      if ((c0) + j < a.Nk) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));                    \
  }
#define AT_BODY_SUM(v, c0)                                                                         \
  if ((c0) + 16 <= a.Nk) {                                                                         \
    _Pragma("unroll") for (int j = 0; j < 16; j++) s4[j & 3] += ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -moff)); \
  } else {                                                                                         \
    _Pragma("unroll") for (int j = 0; j < 16; j++)                                                 \
      if ((c0) + j < a.Nk) s4[j & 3] += ex2_approx(fmaf(__uint_as_float(v[j]), sl2, -moff));       \
  }
  // P = 2^(s*scale*log2e - eoff) as bf16 -> K-major SWIZZLE_128B staging tile; ACC: also the row sum over the ROUNDED values
#define AT_BODY_P(v, c0, ACC)                                                                      \
  {                                                                                                \
    uint32_t pk[8];                                                                                \
    const bool cfull = (c0) + 16 <= a.Nk;                                                          \
    _Pragma("unroll") for (int j = 0; j < 8; j++) {                                                \
      float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), sl2, -eoff));                          \
      float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), sl2, -eoff));                      \
      if (!cfull) {                                                                                \
        p0 = (c0) + 2 * j < a.Nk ? p0 : 0.f;                                                       \
        p1 = (c0) + 2 * j + 1 < a.Nk ? p1 : 0.f;                                                   \
      }                                                                                            \
      __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);                                           \
      if (ACC) {                                                                                   \
        const float2 back = __bfloat1622float2(h2);                                                \
        s4[j & 3] += back.x + back.y;                                                              \
      }                                                                                            \
      pk[j] = *reinterpret_cast<uint32_t*>(&h2);                                                   \
    }                                                                                              \
    _Pragma("unroll") for (int g = 0; g < 2; g++) {                                                \
      const int col = (c0) + g * 8;                                                                \
      const uint32_t kb = (uint32_t)col >> 6, ch = ((uint32_t)col & 63u) >> 3;                     \
      st_shared_v4(rowad + kb * 16384 + ((ch ^ sw) << 4), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]); \
    }                                                                                              \
  }
#define AT_BODY_P_ACC(v, c0) AT_BODY_P(v, c0, true)
#define AT_BODY_P_NOACC(v, c0) AT_BODY_P(v, c0, false)
        // ---- pass 1: row maximum
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
        AT_WALK(AT_BODY_MAX)
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sRed + 4u * (part * 128 + r)), "f"(mx) : "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_max) : "memory");
        float m = -INFINITY;
#pragma unroll
        for (int pp = 0; pp < AT_NP; pp++) {
          float mpart;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mpart) : "r"(sRed + 4u * (pp * 128 + r)));
          m = fmaxf(m, mpart);
        }
        const float moff = m * sl2;
        const uint32_t rowad = sP + (uint32_t)r * 128;
        const uint32_t sw = (uint32_t)r & 7u;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        float eoff = moff;
        if (a.store_p) {
          // ---- training, pass 2: row sum of p = 2^(s*scale*log2e - m*scale*log2e) only (one FFMA + one MUFU.EX2 + one FADD per
          // element); pass 3 recomputes the exponential with the normaliser folded into the offset and writes the NORMALISED bf16
          // probabilities once - instead of storing unnormalised values and rescaling them in place (a shared-memory round trip
          // with two conversions per element).
          AT_WALK(AT_BODY_SUM)
        } else {
          // ---- inference / recompute backward, pass 2: unnormalised bf16 P -> shared memory + row sum over the rounded values
          // (what the P V MMA sees); the 64 output columns are scaled by 1 / sum in the epilogue instead
          AT_WALK(AT_BODY_P_ACC)
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sRed + AT_RED2 + 4u * (part * 128 + r)), "f"(sum) : "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_sum) : "memory");
        float lsum = 0.f;
#pragma unroll
        for (int pp = 0; pp < AT_NP; pp++) {
          float lpart;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lpart) : "r"(sRed + AT_RED2 + 4u * (pp * 128 + r)));
          lsum += lpart;
        }
        o_scale = 1.f;
        if (a.store_p) {
          // the TMA store of the previous tile's P must have finished READING the staging tile before it is overwritten
          if (i > 0) mbar_wait(p_free, (uint32_t)(i - 1) & 1u);
          eoff = moff + __log2f(lsum);   // p / sum = 2^(s*scale*log2e - eoff)
          AT_WALK(AT_BODY_P_NOACC)
        } else {
          o_scale = 1.f / lsum;
        }
        // keys beyond Nkv: P = 0 (their V rows are zero-filled, but 0 * garbage could be NaN)
#pragma unroll
        for (int k = 0; k < 10; k++) {
          const int col = colb + 8 * k;
          if (col >= ((a.Nk + 15) & ~15)) {
            const uint32_t kb = (uint32_t)col >> 6, ch = ((uint32_t)col & 63u) >> 3;
            st_shared_v4(rowad + kb * 16384 + ((ch ^ sw) << 4), 0u, 0u, 0u, 0u);
          }
        }
        if (a.lse && part == 0 && q0 + r < a.N)
          a.lse[bh * a.N + q0 + r] = m * (sl2 * 0.69314718055994531f) + logf(lsum);
        tc_fence_before();
        fence_async_smem();          // generic-proxy writes of P fenced for the async proxy (MMA operand fetch, TMA store)
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);   // AT_SW arrivals complete the phase
      }
      // ---- epilogue: O (normalised here when P was not) TMEM -> bf16 -> global; each thread of the pair takes 32 of the 64 columns
      mbar_wait(o_full, (uint32_t)i & 1u);
      tc_fence_after();
      {
        constexpr int OC = AT_D / AT_NP;   // output columns of this warp (16 with four parts)
        uint32_t v[OC];
        tmem_ld_n<OC>(t_row + AT_O_COL + (uint32_t)(part * OC), v);
        tmem_wait_ld();
        if (q0 + r < a.N) {
          bf16* dst = a.o + ((long)b * a.N + q0 + r) * a.ldo + h * AT_D + part * OC;
#pragma unroll
          for (int g = 0; g < OC / 8; g++) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = __uint_as_float(v[g * 8 + j]) * (MODE == 0 ? o_scale : 1.f);
            store8(dst + g * 8, f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
    }
    if (MODE == 1 && threadIdx.x == 64) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

CMX_API int cmx_attn_fwd(const void* q, int64_t ldq, const void* kv, int64_t ldkv, void* o, int64_t ldo, void* p_out, int64_t ldp,
                         float* lse, int B, int N, int Nk, int heads, float scale, int64_t kv_sample_rows, void* stream) {
  CMX_REQUIRE(q && kv && o, "attn_fwd: null operand");
  if (kv_sample_rows <= 0) kv_sample_rows = Nk;   // key rows per sample in kv (> Nk when this call covers one key chunk only)
  CMX_REQUIRE(Nk >= 1 && Nk <= AT_NK, "attn_fwd: Nkv=%d unsupported (max %d) - use the unfused path", Nk, AT_NK);
  CMX_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && (!p_out || ldp % 8 == 0), "attn_fwd: leading dims must be multiples of 8");
  CMX_REQUIRE(ldq >= heads * AT_D && ldkv >= 2 * heads * AT_D, "attn_fwd: head_dim must be 64");
  CMX_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)kv & 15) == 0 && ((uintptr_t)o & 15) == 0 && ((uintptr_t)p_out & 15) == 0,
              "attn_fwd: pointers must be 16-byte aligned");
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmKV, tmP;
  memset(&tmP, 0, sizeof(tmP));
  int rc = cmx_make_map3(&tmQ, q, (uint64_t)heads * AT_D, (uint64_t)N, (uint64_t)B, (uint64_t)ldq, (uint64_t)N * ldq, AT_D, AT_BM);
  if (rc) return rc;
  rc = cmx_make_map3(&tmKV, kv, (uint64_t)2 * heads * AT_D, (uint64_t)Nk, (uint64_t)B, (uint64_t)ldkv, (uint64_t)kv_sample_rows * ldkv, AT_D, 160);
  if (rc) return rc;
  if (p_out) {
    rc = cmx_make_map3(&tmP, p_out, (uint64_t)Nk, (uint64_t)N, (uint64_t)B * heads, (uint64_t)ldp, (uint64_t)N * ldp, 64, AT_BM);
    if (rc) return rc;
  }
  AttnArgs a;
  a.o = (bf16*)o; a.ldo = ldo; a.lse = lse; a.B = B; a.N = N; a.Nk = Nk; a.heads = heads;
  a.tiles_per_bh = cdiv(N, AT_BM);
  a.total_tiles = (long)a.tiles_per_bh * B * heads;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.scale = scale;
  a.store_p = p_out ? 1 : 0;
  static thread_local PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
    attr_once.mark();
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long grid = a.total_tiles < sms ? a.total_tiles : sms;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(AT_THREADS);
    cfg.dynamicSmemBytes = AT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cmx_use_pdl() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, attn_kernel<0>, tmQ, tmKV, tmP, tmP, a);
  }
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_fwd_kernel");
  return 0;
}

CMX_API int cmx_attn_bwd(const void* d_o, int64_t lddo, const void* kv, int64_t ldkv, const void* p, int64_t ldp, void* ds_out,
                         int64_t ldds, void* dq, int64_t lddq, int B, int N, int Nk, int heads, float scale, void* stream) {
  CMX_REQUIRE(d_o && kv && p && ds_out && dq, "attn_bwd: null operand");
  CMX_REQUIRE(Nk >= 1 && Nk <= AT_NK, "attn_bwd: Nkv=%d unsupported (max %d) - use the unfused path", Nk, AT_NK);
  CMX_REQUIRE(lddo % 8 == 0 && ldkv % 8 == 0 && ldp % 8 == 0 && ldds % 8 == 0 && lddq % 8 == 0, "attn_bwd: leading dims %% 8");
  CMX_REQUIRE(((uintptr_t)d_o & 15) == 0 && ((uintptr_t)kv & 15) == 0 && ((uintptr_t)p & 15) == 0 && ((uintptr_t)ds_out & 15) == 0 &&
                  ((uintptr_t)dq & 15) == 0, "attn_bwd: pointers must be 16-byte aligned");
  if (B == 0 || N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmQ, tmKV, tmP, tmDS;
  int rc = cmx_make_map3(&tmQ, d_o, (uint64_t)heads * AT_D, (uint64_t)N, (uint64_t)B, (uint64_t)lddo, (uint64_t)N * lddo, AT_D, AT_BM);
  if (rc) return rc;
  rc = cmx_make_map3(&tmKV, kv, (uint64_t)2 * heads * AT_D, (uint64_t)Nk, (uint64_t)B, (uint64_t)ldkv, (uint64_t)Nk * ldkv, AT_D, 160);
  if (rc) return rc;
  rc = cmx_make_map3(&tmP, p, (uint64_t)Nk, (uint64_t)N, (uint64_t)B * heads, (uint64_t)ldp, (uint64_t)N * ldp, 64, AT_BM);
  if (rc) return rc;
  rc = cmx_make_map3(&tmDS, ds_out, (uint64_t)Nk, (uint64_t)N, (uint64_t)B * heads, (uint64_t)ldds, (uint64_t)N * ldds, 64, AT_BM);
  if (rc) return rc;
  AttnArgs a;
  a.o = (bf16*)dq; a.ldo = lddq; a.lse = nullptr; a.B = B; a.N = N; a.Nk = Nk; a.heads = heads;
  a.tiles_per_bh = cdiv(N, AT_BM);
  a.total_tiles = (long)a.tiles_per_bh * B * heads;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.scale = scale;
  a.store_p = 0;
  static thread_local PerDeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute(attn bwd): %s", cudaGetErrorString(e));
    attr_once.mark();
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long grid = a.total_tiles < sms ? a.total_tiles : sms;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(AT_THREADS);
    cfg.dynamicSmemBytes = AT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cmx_use_pdl() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, attn_kernel<1>, tmQ, tmKV, tmP, tmDS, a);
  }
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("attn_bwd_kernel");
  return 0;
}
