// GEMM for the CMX hot path on sm_100a.
//
//  * gemm_tc_kernel  — tcgen05.mma (UMMA 128xBNx16, bf16 -> fp32 in TMEM), operands staged by TMA
//    (cp.async.bulk.tensor, SWIZZLE_128B) through an mbarrier ring, warp-specialised:
//    warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2-5 = epilogue
//    (tcgen05.ld 32x32b -> bias/act/DropPath-scale/residual -> vectorised stores or fp32 red.add).
//    Operands may be K-major or MN-major (dgrad uses B MN-major = the untouched [N,K] weight,
//    wgrad uses A and B MN-major = activations as stored), so no transposed copies exist in HBM.
//  * gemm_wmma_kernel — generic strided/batched tensor-core fallback (mma.sync via WMMA) for the
//    shapes the TMA path cannot take (tiny N such as num_classes, odd strides, batched views).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/cmx_b200.h"
#include <cuda.h>
#include <mma.h>
#include <atomic>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

extern std::atomic<long long> g_cmx_launches;

// ------------------------------------------------------------------------------------------------
// Epilogue shared by both kernels
// ------------------------------------------------------------------------------------------------
struct Epi {
  void* C;
  long ldc;
  const float* bias;
  const void* res;
  long ldr;
  const float* row_scale;
  int rows_per_sample;
  int c_dtype, r_dtype, act, atomic;
  float alpha;
  long M, N;
};

__device__ __forceinline__ float epi_scalar(const Epi& e, long row, long col, float acc) {
  float v = acc * e.alpha;
  if (e.bias) v += e.bias[col];
  if (e.act == CMX_ACT_RELU) v = fmaxf(v, 0.f);
  if (e.row_scale) v *= e.row_scale[(int)row / e.rows_per_sample];
  if (e.res) {
    if (e.r_dtype == CMX_F32) v += reinterpret_cast<const float*>(e.res)[row * e.ldr + col];
    else v += __bfloat162float(reinterpret_cast<const bf16*>(e.res)[row * e.ldr + col]);
  }
  return v;
}
__device__ __forceinline__ void epi_store_scalar(const Epi& e, long row, long col, float acc) {
  float v = epi_scalar(e, row, col, acc);
  if (e.c_dtype == CMX_F32) {
    float* p = reinterpret_cast<float*>(e.C) + row * e.ldc + col;
    if (e.atomic) atomicAdd(p, v); else *p = v;
  } else {
    reinterpret_cast<bf16*>(e.C)[row * e.ldc + col] = __float2bfloat16(v);
  }
}
// epilogue math for 8 consecutive columns kept in registers (used by the TMA-store path); sbias = shared-memory
// copy of the bias slice of this tile (zeros without bias).  MODE 0 (alpha = 1, no bias / activation / row scale: every
// dgrad and attention-style product) compiles to nothing; MODE 1 (alpha and/or bias: the forward linears) is four
// packed FFMA2; MODE 2 adds the ReLU clamp and the DropPath row scale.  The single
// epilogue warp per SM sub-partition is issue-latency bound, so instructions per element are what the epilogue costs.
// Out-of-range columns/rows only skip the residual load.
template <int MODE>  // 0 plain, 1 alpha/bias only, 2 also ReLU and/or row scale
__device__ __forceinline__ void epi_math8(const Epi& e, long row, long col, const float (&sbias)[8], bool row_ok, float rs,
                                          float (&v)[8]) {
  if (MODE != 0) {
    const float2 a2 = make_float2(e.alpha, e.alpha);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 t = make_float2(sbias[2 * i], sbias[2 * i + 1]);
      ffma2(t, make_float2(v[2 * i], v[2 * i + 1]), a2);
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
    if (MODE == 2) {
      const float lo = e.act == CMX_ACT_RELU ? 0.f : -INFINITY;   // ReLU as a clamp: no branch in the unrolled loop
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], lo) * rs;
    }
  }
  if (e.res && row_ok) {
    if (col + 8 <= e.N) {
      float r[8];
      if (e.r_dtype == CMX_F32) load8(reinterpret_cast<const float*>(e.res) + row * e.ldr + col, r);
      else load8(reinterpret_cast<const bf16*>(e.res) + row * e.ldr + col, r);
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] += r[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (col + i < e.N)
          v[i] += e.r_dtype == CMX_F32 ? reinterpret_cast<const float*>(e.res)[row * e.ldr + col + i]
                                       : __bfloat162float(reinterpret_cast<const bf16*>(e.res)[row * e.ldr + col + i]);
    }
  }
}

// 8 consecutive columns, all in range, 16B-aligned addresses (checked on the host).  PLAIN: alpha = 1 and no bias /
// activation / row scale (every split-K weight gradient) - the accumulator goes straight to the store / red.add.
template <bool PLAIN = false>
__device__ __forceinline__ void epi_store_vec8(const Epi& e, long row, long col, float* v) {
  if (!PLAIN) {
    const float rs = e.row_scale ? e.row_scale[(int)row / e.rows_per_sample] : 1.f;
    const float lo = e.act == CMX_ACT_RELU ? 0.f : -INFINITY;
    float b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = 0.f;
    if (e.bias) load8(e.bias + col, b);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = fmaxf(fmaf(v[i], e.alpha, b[i]), lo) * rs;
  }
  if (e.res) {
    float r[8];
    if (e.r_dtype == CMX_F32) load8(reinterpret_cast<const float*>(e.res) + row * e.ldr + col, r);
    else load8(reinterpret_cast<const bf16*>(e.res) + row * e.ldr + col, r);
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] += r[i];
  }
  if (e.c_dtype == CMX_F32) {
    float* p = reinterpret_cast<float*>(e.C) + row * e.ldc + col;
    if (e.atomic) {
      atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(v[4], v[5], v[6], v[7]));
    } else {
      store8(p, v);
    }
  } else {
    store8(reinterpret_cast<bf16*>(e.C) + row * e.ldc + col, v);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 GEMM kernel — persistent, warp-specialised.
//   grid  = min(#tiles, #SMs) CTAs, each walking tiles t = blockIdx.x, +gridDim.x, ... (N-tile fastest so the
//           weight tile stays L2/smem-hot); a tile is 128 x BN of one (batch, split-K slice).
//   warp 0 = TMA producer: the shared-memory ring (BK = 64 bf16 = one 128-byte swizzle atom per row) keeps
//            running ACROSS tile boundaries, so the loads of tile i+1 overlap the MMAs/epilogue of tile i.
//   warp 1 = MMA issuer (one elected lane) + TMEM allocator; accumulators are double-buffered in TMEM
//            (2 x BN fp32 columns) so the MMAs of tile i+1 overlap the epilogue of tile i.
//   warps 2.. = epilogue (4 or 8 warps, see tc_epi_warps; with 8, two warps per TMEM lane quarter take interleaved
//            32-column chunks):
//            tcgen05.ld -> bias/act/DropPath-scale/residual -> 16-byte stores or fp32 red.add (split-K).
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
// Two CTA shapes.  BN <= 128: 4 epilogue warps (one per TMEM lane quarter), <= 110 KB of shared memory, 2 x BN (<= 256)
// TMEM columns and <= 168 registers x 192 threads, so TWO CTAs are resident per SM: the epilogue of one overlaps the
// loads/MMAs/epilogue of the other, and kernels of the other streams (the second modality branch, the weight-gradient
// companion streams) or the PDL-launched successor can share the SM.  BN > 128: 8 epilogue warps (two per lane quarter,
// interleaved 32-column chunks), the whole SM.
// EW = epilogue warps per CTA (4 = one per TMEM lane quarter, 8 = two per quarter on interleaved 32-column chunks).  The epilogue
// of the K <= 256 shapes is the whole kernel (ncu, profiles/r2_gemm_epilogue_*: every warp issues one instruction per ~8 cycles,
// the MMA issuer waits for the accumulator buffer all the time), so the two-CTA-per-SM shapes can run 8 epilogue warps each
// (16 per SM = 4 per scheduler) with ONE 4 KB staging buffer per warp instead of two.
__host__ __device__ constexpr int tc_threads(int ew) { return 64 + 32 * ew; }
// CPS = CTAs per SM.  2: the default for BN <= 128 (one CTA's epilogue overlaps the other's loads / MMAs; 110 KB each = a
// 2-deep operand ring).  1: the whole SM for one CTA - required for BN > 128, and an opt-in deep-ring variant of BN = 128
// (CMX_GEMM_DEEP_K, measured slower, see deep_k_env).
__host__ __device__ constexpr int tc_nbuf(int ew, int cps) { return ew == 4 ? 2 : 1; }  // staging buffers per epilogue warp
__host__ __device__ constexpr uint32_t tc_cstage_bytes(int ew, int cps) { return ew * tc_nbuf(ew, cps) * 4096; }  // (32 rows x 128 B) each
__host__ __device__ constexpr int tc_smem_budget(int cps) { return cps == 2 ? 110 * 1024 : 216 * 1024; }

template <int EMODE_, int CDT_, int RES_, bool FULL_>
struct EK {   // compile-time description of one epilogue body, see gemm_tc_kernel
  static constexpr int EMODE = EMODE_, CDT = CDT_, RES = RES_;
  static constexpr bool FULL = FULL_;
};

struct TcSched {
  int tiles_n, tiles_m, batch2, nbatch, splits;
  int kb_total, kb_per, stages;
  long total_tiles;
  long sC1, sC2;  // element strides of C per batch index
  long sBias1, sR1, sS1;  // batch1 strides (elements) of bias / residual / row_scale: grouped (per-branch weights) launches
  int tma_store;  // 0 = per-thread global stores (atomics / odd layouts), 1 = smem-staged TMA bulk store
  long long* trace;  // debug: CTA 0 writes clock64 stamps [role][tile][4] (role 0 producer, 1 mma, 2 epilogue, 3/4 epilogue detail)
};
#define TC_TRACE(role, tile, slot)                                                              \
  do {                                                                                          \
    if (sc.trace && blockIdx.x == 0 && (tile) < 64) sc.trace[((role) * 64 + (tile)) * 4 + (slot)] = clock64(); \
  } while (0)


template <int BN, bool A_MN, bool B_MN, bool BATCHED, int EW, int CPS>
__global__ void __launch_bounds__(tc_threads(EW), CPS) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmC, Epi epi0,
                                                               TcSched sc) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t A_BYTES = TC_BM * TC_BK * 2;
  constexpr uint32_t B_BYTES = BN * TC_BK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t ACC_STRIDE = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // TMEM columns per accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  constexpr int TC_EPI_WARPS = EW;
  constexpr int NH = TC_EPI_WARPS / 4;  // warps per TMEM lane quarter = column interleave factor
  constexpr int NBUF = tc_nbuf(EW, CPS);
  constexpr bool SIMPLE = CPS == 2;  // one register set: the co-resident CTA hides the tcgen05.ld latency
  constexpr uint32_t TC_CSTAGE_BYTES = tc_cstage_bytes(EW, CPS);
  const int stages = sc.stages;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t cstage_base = smem_base + stages * STAGE_BYTES;                   // 1024-aligned (stage sizes are)
  const uint32_t bias_base = cstage_base + (sc.tma_store ? TC_CSTAGE_BYTES : 0u);  // float[2][256]
  const uint32_t bar_base = bias_base + 2048u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t tfull_bar = bar_base + 16u * stages;       // [2]
  const uint32_t tempty_bar = tfull_bar + 16u;              // [2]
  const uint32_t tmem_ptr_addr = tempty_bar + 16u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (sc.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int s = 0; s < stages; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(tfull_bar + 8u * a, 1);
      mbar_init(tempty_bar + 8u * a, TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  pdl_wait();  // PDL: everything above overlapped the previous kernel's tail; its results are visible from here on

  // tile -> (n tile, m tile, batch, split) ; identical arithmetic in all three roles
  auto decode = [&](long tl, int& n0, int& m0, int& b1, int& b2, int& kb_begin, int& nkb, int& split_idx) {
    unsigned t = (unsigned)tl;  // the host guarantees total_tiles < 2^31
    const int tn = (int)(t % (unsigned)sc.tiles_n);
    t /= (unsigned)sc.tiles_n;
    const int tm = (int)(t % (unsigned)sc.tiles_m);
    t /= (unsigned)sc.tiles_m;
    const int bb = (int)(t % (unsigned)sc.nbatch);
    split_idx = (int)(t / (unsigned)sc.nbatch);
    n0 = tn * BN;
    m0 = tm * TC_BM;
    b2 = bb % sc.batch2;
    b1 = bb / sc.batch2;
    kb_begin = split_idx * sc.kb_per;
    int kb_end = kb_begin + sc.kb_per;
    if (kb_end > sc.kb_total) kb_end = sc.kb_total;
    nkb = kb_end - kb_begin;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (long t = blockIdx.x; t < sc.total_tiles; t += gridDim.x) {
        int n0, m0, b1, b2, kb_begin, nkb, sp;
        decode(t, n0, m0, b1, b2, kb_begin, nkb, sp);
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (i == 0) TC_TRACE(0, it, 0);
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          const int k = (kb_begin + i) * TC_BK;
          if (!BATCHED) {
            if (!A_MN) {
              tma_load_2d(sa, &tmA, full_bar(s), k, m0);
            } else {
#pragma unroll
              for (int c = 0; c < TC_BM / 64; c++) tma_load_2d(sa + c * 8192, &tmA, full_bar(s), m0 + c * 64, k);
            }
            if (!B_MN) {
              tma_load_2d(sb, &tmB, full_bar(s), k, n0);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; c++) tma_load_2d(sb + c * 8192, &tmB, full_bar(s), n0 + c * 64, k);
            }
          } else {
            if (!A_MN) {
              tma_load_4d(sa, &tmA, full_bar(s), k, m0, b2, b1);
            } else {
#pragma unroll
              for (int c = 0; c < TC_BM / 64; c++) tma_load_4d(sa + c * 8192, &tmA, full_bar(s), m0 + c * 64, k, b2, b1);
            }
            if (!B_MN) {
              tma_load_4d(sb, &tmB, full_bar(s), k, n0, b2, b1);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; c++) tma_load_4d(sb + c * 8192, &tmB, full_bar(s), n0 + c * 64, k, b2, b1);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t it = 0, lt = 0;
      for (long t = blockIdx.x; t < sc.total_tiles; t += gridDim.x, lt++) {
        int n0, m0, b1, b2, kb_begin, nkb, sp;
        decode(t, n0, m0, b1, b2, kb_begin, nkb, sp);
        const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
        TC_TRACE(1, lt, 0);
        mbar_wait(tempty_bar + 8u * acc, aph ^ 1u);  // epilogue has drained this accumulator buffer
        TC_TRACE(1, lt, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int i = 0; i < nkb; i++, it++) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; MN-major: 16 K-rows of 128 B = 2048 B
            const uint64_t da = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
            tc_mma_bf16(d_tmem, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          if (i == 0) TC_TRACE(1, lt, 2);
          tc_commit(empty_bar(s));  // frees the smem slot once the MMAs that read it retire
        }
        tc_commit(tfull_bar + 8u * acc);
        TC_TRACE(1, lt, 3);
      }
    }
  } else {
    // ================= epilogue: warps 2..; TMEM lane quarter = warp % 4, NH warps per quarter =================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;  // 0 .. 32 * TC_EPI_WARPS - 1
    const uint32_t my_stage = cstage_base + (uint32_t)(warp - 2) * (NBUF * 4096u);
    uint32_t lt = 0, nstore = 0;
    float bias_pf[BN / 32];  // next tile's bias values (lane-strided), NH == 1 only
#pragma unroll
    for (int i = 0; i < BN / 32; i++) bias_pf[i] = 0.f;
    for (long t = blockIdx.x; t < sc.total_tiles; t += gridDim.x, lt++) {
      int n0, m0, b1, b2, kb_begin, nkb, sp;
      decode(t, n0, m0, b1, b2, kb_begin, nkb, sp);
      Epi epi = epi0;
      if (BATCHED) {
        const long off = (long)b1 * sc.sC1 + (long)b2 * sc.sC2;
        if (epi.c_dtype == CMX_F32) epi.C = reinterpret_cast<float*>(epi.C) + off;
        else epi.C = reinterpret_cast<bf16*>(epi.C) + off;
        if (epi.bias) epi.bias += (long)b1 * sc.sBias1;
        if (epi.row_scale) epi.row_scale += (long)b1 * sc.sS1;
        if (epi.res) {
          if (epi.r_dtype == CMX_F32) epi.res = reinterpret_cast<const float*>(epi.res) + (long)b1 * sc.sR1;
          else epi.res = reinterpret_cast<const bf16*>(epi.res) + (long)b1 * sc.sR1;
        }
      }
      if (sp != 0) {  // split-K: bias / residual are contributed once, by slice 0
        epi.bias = nullptr;
        epi.res = nullptr;
      }
      uint32_t sb_addr;
      if constexpr (NH == 1) {
        // bias slice of this tile -> a WARP-PRIVATE shared-memory copy (no CTA barrier); the values were fetched into
        // registers one tile ahead, so the global-load latency hides behind the previous tile's epilogue
        sb_addr = bias_base + (uint32_t)(warp - 2) * (BN * 4u);
        if (lt == 0 && epi0.bias) {
          const float* bp = epi0.bias + (BATCHED ? (long)b1 * sc.sBias1 : 0l);
#pragma unroll
          for (int i = 0; i < BN / 32; i++) bias_pf[i] = (n0 + lane + 32 * i < epi0.N) ? bp[n0 + lane + 32 * i] : 0.f;
        }
        {
          // (the copy is read unconditionally by the alpha / bias epilogue bodies: zeros without a bias, or for split-K
          // slices other than the first)
          __syncwarp();   // every lane has finished reading the previous tile's copy
#pragma unroll
          for (int i = 0; i < BN / 32; i++)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb_addr + 4u * (lane + 32 * i)), "f"(epi.bias ? bias_pf[i] : 0.f) : "memory");
          __syncwarp();
        }
        const long tn = t + gridDim.x;
        if (epi0.bias && tn < sc.total_tiles) {
          int n1, m1, c1, c2, kb1, nk1, sp1;
          decode(tn, n1, m1, c1, c2, kb1, nk1, sp1);
          const float* bp = epi0.bias + (BATCHED ? (long)c1 * sc.sBias1 : 0l);
#pragma unroll
          for (int i = 0; i < BN / 32; i++) bias_pf[i] = (n1 + lane + 32 * i < epi0.N) ? bp[n1 + lane + 32 * i] : 0.f;
        }
        if (warp == 4 && lane == 0) { TC_TRACE(2, lt, 0); TC_TRACE(2, lt, 1); }
      } else {
        // bias slice of this tile -> shared memory (double-buffered by tile parity; one named barrier per tile)
        sb_addr = bias_base + (lt & 1u) * 1024u;
        {   // (read unconditionally by the alpha / bias epilogue bodies: zeros without a bias)
          for (int i = etid; i < BN; i += 32 * TC_EPI_WARPS) {
            const float bv = (epi.bias && n0 + i < epi.N) ? epi.bias[n0 + i] : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb_addr + 4u * i), "f"(bv) : "memory");
          }
        }
        if (warp == 4 && lane == 0) TC_TRACE(2, lt, 0);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        if (warp == 4 && lane == 0) TC_TRACE(2, lt, 1);
      }
      const uint32_t acc = lt & 1u, aph = (lt >> 1) & 1u;
      if (warp == 4 && lane == 0) TC_TRACE(2, lt, 1);
      mbar_wait(tfull_bar + 8u * acc, aph);
      if (warp == 4 && lane == 0) TC_TRACE(2, lt, 2);
      tc_fence_after();
      const long row = (long)m0 + q * 32 + lane;
      const bool row_ok = row < epi.M;
      const float rs = (epi.row_scale && row_ok) ? epi.row_scale[(int)row / epi.rows_per_sample] : 1.f;
      const uint32_t t_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      if (sc.tma_store) {
        // ---- smem-staged TMA store: warp-private [32 rows x 128 B] boxes, SWIZZLE_128B, double buffered.
        // The tcgen05.ld of chunk k+1 is issued before the math / st.shared of chunk k (two register sets).
        const int UC = epi.c_dtype == CMX_F32 ? 1 : 2;   // chunks per store unit: 32 fp32 or 64 bf16 columns = 128 B
        // The epilogue is the critical path of every K <= 512 shape (ncu: one instruction per ~8 cycles and warp, a fifth of
        // them branches), so the three hot combinations are compiled as branch-free bodies, chosen per tile by a warp-uniform
        // jump; everything else (ragged tiles, other dtype / residual combinations) takes the generic body.
        //   EK<EMODE, CDT, RES, FULL>: EMODE 0 plain / 1 alpha+bias / 2 also ReLU + row scale; CDT 0 bf16, 1 fp32, 2 runtime;
        //   RES 0 none, 1 fp32, 2 runtime; FULL: all 128 rows and BN columns of the tile are in range
        auto process = [&](auto kind_tag, uint32_t* r, int c) {
          using K = decltype(kind_tag);
          constexpr int EMODE = K::EMODE;
          const bool f32out = K::CDT == 2 ? (epi.c_dtype == CMX_F32) : (K::CDT == 1);
          const int UCk = f32out ? 1 : 2;
          const int cc = c % UCk;
          const uint32_t buf = my_stage + (NBUF == 2 ? (nstore & 1u) * 4096u : 0u);
          const bool trc = (warp == 4 && lane == 0);
          if (cc == 0) {
            if (trc) TC_TRACE(3, lt, 0);
            if (lane == 0) tma_store_wait_read<NBUF - 1>();  // the store that last used this buffer has read it
            __syncwarp();
            if (trc) TC_TRACE(3, lt, 1);
          }
          // residual row segment of this chunk: one 64-bit address computation per chunk
          const long ccol = (long)n0 + c * 32;
          const bool cfull = K::FULL || ccol + 32 <= epi.N;
          const bool rres = K::RES == 0 ? false : (K::FULL && K::RES == 1 ? true : (epi.res != nullptr && row_ok));
          const bool res_f32 = K::RES == 1 || epi.r_dtype == CMX_F32;
          const float* res32 = reinterpret_cast<const float*>(epi.res) + row * epi.ldr + ccol;
          const bf16* res16 = reinterpret_cast<const bf16*>(epi.res) + row * epi.ldr + ccol;
          float rr[2][8];   // residual values, fetched one 8-column group ahead of the math
          auto res_fetch = [&](int g) {
            if (res_f32) load8(res32 + g * 8, rr[g & 1]);
            else load8(res16 + g * 8, rr[g & 1]);
          };
          if (rres && cfull) res_fetch(0);
#pragma unroll
          for (int g = 0; g < 4; g++) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = __uint_as_float(r[g * 8 + j]);
            if (EMODE != 0) {
              float sbv[8];
              const uint32_t ba = sb_addr + 4u * (c * 32 + g * 8);
              asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sbv[0]), "=f"(sbv[1]), "=f"(sbv[2]), "=f"(sbv[3]) : "r"(ba));
              asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sbv[4]), "=f"(sbv[5]), "=f"(sbv[6]), "=f"(sbv[7]) : "r"(ba + 16u));
              const float2 a2 = make_float2(epi.alpha, epi.alpha);
#pragma unroll
              for (int i = 0; i < 4; i++) {
                float2 t = make_float2(sbv[2 * i], sbv[2 * i + 1]);
                ffma2(t, make_float2(v[2 * i], v[2 * i + 1]), a2);
                v[2 * i] = t.x;
                v[2 * i + 1] = t.y;
              }
              if (EMODE == 2) {
                const float lo = epi.act == CMX_ACT_RELU ? 0.f : -INFINITY;   // ReLU as a clamp: no branch in the unrolled loop
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = fmaxf(v[i], lo) * rs;
              }
            }
            if (rres) {
              if (cfull) {
                if (g < 3) res_fetch(g + 1);
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] += rr[g & 1][i];
              } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
                  if (ccol + g * 8 + i < epi.N)
                    v[i] += res_f32 ? res32[g * 8 + i] : __bfloat162float(res16[g * 8 + i]);
              }
            }
            const uint32_t sw = (uint32_t)(lane & 7);
            if (f32out) {
              const uint32_t j0 = (uint32_t)(g * 2);
              st_shared_v4(buf + lane * 128u + ((j0 ^ sw) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
                           __float_as_uint(v[2]), __float_as_uint(v[3]));
              st_shared_v4(buf + lane * 128u + (((j0 + 1) ^ sw) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]),
                           __float_as_uint(v[6]), __float_as_uint(v[7]));
            } else {
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; j++) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                pk[j] = *reinterpret_cast<uint32_t*>(&h2);
              }
              const uint32_t j0 = (uint32_t)(cc * 4 + g);
              st_shared_v4(buf + lane * 128u + ((j0 ^ sw) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
          }
          // last chunk of the unit (or the unit is cut short by N): publish the box
          const bool unit_done = (cc == UCk - 1) || (!K::FULL && n0 + (c + 1) * 32 >= epi.N);
          if (trc) TC_TRACE(3, lt, 2 + (cc ? 1 : 0));
          if (unit_done) {
            fence_async_smem();
            __syncwarp();
            if (trc) TC_TRACE(4, lt, 0);
            if (lane == 0) {
              const int ucol0 = n0 + (c - cc) * 32;
              if (BATCHED) tma_store_4d(&tmC, buf, ucol0, m0 + q * 32, b2, b1);
              else tma_store_2d(&tmC, buf, ucol0, m0 + q * 32);
              tma_store_commit();
            }
            if (trc) TC_TRACE(4, lt, 1);
            nstore++;
          }
        };
        // k-th chunk of this warp (units half, half+NH, ...), -1 when exhausted; a trailing partial unit
        // (BN = 160 with bf16) is excluded here and handled by the per-thread store path below
        auto chunk_of = [&](auto kind_tag, int k) -> int {
          using K = decltype(kind_tag);
          const int UCk = K::CDT == 2 ? UC : (K::CDT == 1 ? 1 : 2);
          const int u = half + NH * (k / UCk);
          const int c = u * UCk + (k % UCk);
          if ((u + 1) * UCk * 32 > BN || (!K::FULL && n0 + c * 32 >= epi.N)) return -1;
          return c;
        };
        auto run_chunks = [&](auto kind_tag) {
          if constexpr (SIMPLE) {
            // two resident CTAs per SM hide the tcgen05.ld latency of each other: one register set, no spills
            uint32_t ra[32];
#pragma unroll 1
            for (int k = 0;; k++) {
              const int c = chunk_of(kind_tag, k);
              if (c < 0) break;
              tmem_ld32(t_addr + (uint32_t)(c * 32), ra);
              tmem_wait_ld_dep(ra);
              process(kind_tag, ra, c);
            }
            return;
          }
          uint32_t ra[32], rb[32];
          int cA = chunk_of(kind_tag, 0);
          if (cA >= 0) tmem_ld32(t_addr + (uint32_t)(cA * 32), ra);
#pragma unroll 1
          for (int k = 0; cA >= 0; k += 2) {
            tmem_wait_ld_dep(ra);
            const int cB = chunk_of(kind_tag, k + 1);
            if (cB >= 0) tmem_ld32(t_addr + (uint32_t)(cB * 32), rb);
            process(kind_tag, ra, cA);
            if (cB < 0) break;
            tmem_wait_ld_dep(rb);
            cA = chunk_of(kind_tag, k + 2);
            if (cA >= 0) tmem_ld32(t_addr + (uint32_t)(cA * 32), ra);
            process(kind_tag, rb, cB);
          }
        };
        // warp-uniform choice between separately compiled epilogue bodies (a real branch, not predication)
        const bool m2 = epi.act != CMX_ACT_NONE || epi.row_scale, m1 = epi.alpha != 1.f || epi.bias;
        if constexpr (SIMPLE) {
          const bool full = m0 + TC_BM <= epi.M && n0 + BN <= epi.N;
          if (full && !m1 && !m2 && !epi.res && epi.c_dtype == CMX_BF16) run_chunks(EK<0, 0, 0, true>{});        // data gradients
          else if (full && !m2 && !epi.res && epi.c_dtype == CMX_BF16) run_chunks(EK<1, 0, 0, true>{});           // x W^T + b
          else if (full && epi.res && epi.r_dtype == CMX_F32 && epi.c_dtype == CMX_F32) run_chunks(EK<2, 1, 1, true>{});  // residual stream
          else run_chunks(EK<2, 2, 2, false>{});
        } else {  // whole-SM shapes (off the default policy)
          run_chunks(EK<2, 2, 2, false>{});
        }
        if ((BN / 32) % UC != 0) {
          // BN = 160 with 64-column bf16 boxes leaves a 32-column remainder: a full box would spill into the
          // neighbouring tile, so that chunk takes the per-thread store path (done by the warp owning that unit)
          const int c = BN / 32 - 1;
          const int u = c / UC;
          if ((u % NH) == half && n0 + c * 32 < epi.N) {
            uint32_t r[32];
            tmem_ld32(t_addr + (uint32_t)(c * 32), r);
            tmem_wait_ld();
            if (row_ok) {
#pragma unroll
              for (int g = 0; g < 4; g++) {
                const long col = (long)n0 + c * 32 + g * 8;
                if (col + 8 <= epi.N) {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; j++) v[j] = __uint_as_float(r[g * 8 + j]);
                  epi_store_vec8(epi, row, col, v);
                } else if (col < epi.N) {
#pragma unroll
                  for (int j = 0; j < 8; j++)
                    if (col + j < epi.N) epi_store_scalar(epi, row, col + j, __uint_as_float(r[g * 8 + j]));
                }
              }
            }
          }
        }
      } else {
        auto run_direct = [&](auto plain_tag) {
          constexpr bool PLAIN = decltype(plain_tag)::value;
#pragma unroll 1
          for (int c = half; c < BN / 32; c += NH) {
            if (n0 + c * 32 >= epi.N) break;
            uint32_t r[32];
            tmem_ld32(t_addr + (uint32_t)(c * 32), r);
            tmem_wait_ld();
            if (row_ok) {
#pragma unroll
              for (int g = 0; g < 4; g++) {
                const long col = (long)n0 + c * 32 + g * 8;
                if (col + 8 <= epi.N) {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; j++) v[j] = __uint_as_float(r[g * 8 + j]);
                  epi_store_vec8<PLAIN>(epi, row, col, v);
                } else if (col < epi.N) {  // ragged tail (e.g. Nkv = 300)
#pragma unroll
                  for (int j = 0; j < 8; j++)
                    if (col + j < epi.N) epi_store_scalar(epi, row, col + j, __uint_as_float(r[g * 8 + j]));
                }
              }
            }
          }
        };
        if (epi.alpha == 1.f && !epi.bias && epi.act == CMX_ACT_NONE && !epi.row_scale) run_direct(std::true_type{});
        else run_direct(std::false_type{});
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8u * acc);
      if (warp == 4 && lane == 0) TC_TRACE(2, lt, 3);
    }
    if (sc.tma_store && lane == 0) tma_store_wait_all();  // global writes complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Generic fallback: 64x64x32 tiles, 4 warps, WMMA bf16 16x16x16, arbitrary strides + 2-level batch.
// ------------------------------------------------------------------------------------------------
struct GenArgs {
  const bf16* A;
  const bf16* B;
  long sAm, sAk, sBk, sBn;  // element strides of A(m,k) and B(k,n)
  long K;
  int batch2;
  long sA1, sA2, sB1, sB2, sC1, sC2;
  long sBias1, sR1, sS1;
  long k_per_split;
};

constexpr int GB = 64, GK = 32;

__global__ void __launch_bounds__(128) gemm_wmma_kernel(GenArgs g, Epi epi) {
  pdl_trigger();
  using namespace nvcuda;
  __shared__ __align__(32) bf16 As[GB][GK + 8];
  __shared__ __align__(32) bf16 Bs[GK][GB + 8];
  __shared__ __align__(32) float Cs[GB][GB + 4];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const long m0 = (long)blockIdx.y * GB, n0 = (long)blockIdx.x * GB;
  int zb = blockIdx.z;
  const int nsplit = (int)((g.K + g.k_per_split - 1) / g.k_per_split);
  const int split = zb % nsplit;
  zb /= nsplit;
  const int b2 = zb % g.batch2, b1 = zb / g.batch2;
  const bf16* A = g.A + b1 * g.sA1 + b2 * g.sA2;
  const bf16* B = g.B + b1 * g.sB1 + b2 * g.sB2;
  Epi e = epi;
  if (e.c_dtype == CMX_F32) e.C = reinterpret_cast<float*>(e.C) + b1 * g.sC1 + b2 * g.sC2;
  else e.C = reinterpret_cast<bf16*>(e.C) + b1 * g.sC1 + b2 * g.sC2;
  if (e.bias) e.bias += b1 * g.sBias1;
  if (e.row_scale) e.row_scale += b1 * g.sS1;
  if (e.res) {
    if (e.r_dtype == CMX_F32) e.res = reinterpret_cast<const float*>(e.res) + b1 * g.sR1;
    else e.res = reinterpret_cast<const bf16*>(e.res) + b1 * g.sR1;
  }
  if (split != 0) { e.bias = nullptr; e.res = nullptr; }

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) wmma::fill_fragment(acc[i][j], 0.f);

  const long k_begin = (long)split * g.k_per_split;
  long k_end = k_begin + g.k_per_split;
  if (k_end > g.K) k_end = g.K;
  const bool a_kfast = (g.sAk == 1);  // iterate the contiguous index fastest for coalescing
  const bool b_nfast = (g.sBn == 1);
  const bf16 zero = __float2bfloat16(0.f);

  for (long k0 = k_begin; k0 < k_end; k0 += GK) {
    for (int i = tid; i < GB * GK; i += 128) {
      int m, k;
      if (a_kfast) { m = i / GK; k = i % GK; } else { k = i / GB; m = i % GB; }
      const long gm = m0 + m, gk = k0 + k;
      As[m][k] = (gm < epi.M && gk < k_end) ? A[gm * g.sAm + gk * g.sAk] : zero;
    }
    for (int i = tid; i < GK * GB; i += 128) {
      int k, n;
      if (b_nfast) { k = i / GB; n = i % GB; } else { n = i / GK; k = i % GK; }
      const long gk = k0 + k, gn = n0 + n;
      Bs[k][n] = (gn < epi.N && gk < k_end) ? B[gk * g.sBk + gn * g.sBn] : zero;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> fa[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb[2];
#pragma unroll
      for (int i = 0; i < 2; i++) wmma::load_matrix_sync(fa[i], &As[wm * 32 + i * 16][kk], GK + 8);
#pragma unroll
      for (int j = 0; j < 2; j++) wmma::load_matrix_sync(fb[j], &Bs[kk][wn * 32 + j * 16], GB + 8);
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
      wmma::store_matrix_sync(&Cs[wm * 32 + i * 16][wn * 32 + j * 16], acc[i][j], GB + 4, wmma::mem_row_major);
  __syncthreads();
  for (int i = tid; i < GB * GB; i += 128) {
    const int m = i / GB, n = i % GB;
    const long gm = m0 + m, gn = n0 + n;
    if (gm < epi.M && gn < epi.N) epi_store_scalar(e, gm, gn, Cs[m][n]);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: dim0 = contiguous extent, dim1 = strided extent (row stride ld elements)
static int make_map(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint32_t box0,
                    uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) CMX_FAIL(-2, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CMX_FAIL(-3, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu ld %llu box %u x %u", (int)r,
                                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)ld, box0, box1);
  return 0;
}

// 4-D bf16 tensor map for batched operands: (dim0 contiguous, dim1 rows with stride ld, batch2, batch1)
static int make_map4(CUtensorMap* tm, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t ld, uint64_t nb2, uint64_t s2,
                     uint64_t nb1, uint64_t s1, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) CMX_FAIL(-2, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {dim0, dim1, nb2, nb1};
  // a batch dimension of extent 1 may carry stride 0 from the caller; TMA wants a positive 16-byte multiple
  cuuint64_t strides[3] = {ld * 2, (nb2 > 1 ? s2 : ld) * 2, (nb1 > 1 ? s1 : ld) * 2};
  cuuint32_t box[4] = {box0, box1, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    CMX_FAIL(-3, "cuTensorMapEncodeTiled(4d) failed (%d): dims %llu x %llu x %llu x %llu strides %llu %llu %llu", (int)r,
             (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)nb2, (unsigned long long)nb1,
             (unsigned long long)strides[0], (unsigned long long)strides[1], (unsigned long long)strides[2]);
  return 0;
}

// C tensor map for the TMA-store epilogue: box = 128 bytes x 32 rows, SWIZZLE_128B; rank 2 or 4 (batched)
static int make_map_c(CUtensorMap* tm, const CmxGemm* g, bool batched) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) CMX_FAIL(-2, "cuTensorMapEncodeTiled unavailable");
  const bool f32 = g->c_dtype == CMX_F32;
  const uint64_t es_ = f32 ? 4 : 2;
  cuuint64_t dims[4] = {(cuuint64_t)g->N, (cuuint64_t)g->M, (cuuint64_t)g->batch2, (cuuint64_t)g->batch1};
  cuuint64_t strides[3] = {(cuuint64_t)g->ldc * es_, (cuuint64_t)(g->batch2 > 1 ? g->sC2 : g->ldc) * es_,
                           (cuuint64_t)(g->batch1 > 1 ? g->sC1 : g->ldc) * es_};
  cuuint32_t box[4] = {f32 ? 32u : 64u, 32u, 1u, 1u};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, batched ? 4 : 2, g->C, dims,
                   strides, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CMX_FAIL(-3, "cuTensorMapEncodeTiled(C) failed (%d)", (int)r);
  return 0;
}

static long long* g_tc_trace = nullptr;
CMX_API int cmx_debug_set_gemm_trace(void* buf) {  // debug only: device buffer of 3*64*4 int64, or NULL
  g_tc_trace = reinterpret_cast<long long*>(buf);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static bool tc_eligible(const CmxGemm* g) {
  const bool batched = g->batch1 != 1 || g->batch2 != 1;
  if (g->M < 1 || g->N < 1 || g->K < 1) return false;
  if ((g->lda % 8) || (g->ldb % 8) || (g->ldc % 8)) return false;  // 16-byte row strides (TMA) / vector epilogue
  if (!aligned16(g->A) || !aligned16(g->B) || !aligned16(g->C)) return false;
  if (g->residual && ((g->ldr % 8) || !aligned16(g->residual))) return false;
  if (g->bias && !aligned16(g->bias)) return false;
  if (g->trans_a && !g->trans_b) return false;  // (MN, K) combination not instantiated
  if ((g->split_k > 1 || g->accumulate) && g->c_dtype != CMX_F32) return false;
  if (g->split_k > 1 && g->act != CMX_ACT_NONE) return false;  // non-linear epilogue needs the full sum
  if (batched) {
    if ((g->residual || g->bias || g->row_scale) && g->batch2 != 1) return false;  // epilogue operands follow batch1 only
    if ((g->sBias1 % 4) || (g->sR1 % 8)) return false;
    if ((g->sA1 % 8) || (g->sA2 % 8) || (g->sB1 % 8) || (g->sB2 % 8) || (g->sC1 % 8) || (g->sC2 % 8)) return false;
    if ((g->batch1 > 1 && (g->sA1 <= 0 || g->sB1 <= 0)) || (g->batch2 > 1 && (g->sA2 <= 0 || g->sB2 <= 0))) return false;
  }
  return true;
}

template <int BN, bool A_MN, bool B_MN, bool BATCHED, int EW, int CPS>
static int launch_tc(const CmxGemm* g, const Epi& epi, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  int rc;
  if (!BATCHED) {
    if (!A_MN) rc = make_map(&tmA, g->A, g->K, g->M, g->lda, TC_BK, TC_BM);
    else rc = make_map(&tmA, g->A, g->M, g->K, g->lda, 64, TC_BK);
    if (rc) return rc;
    if (!B_MN) rc = make_map(&tmB, g->B, g->K, g->N, g->ldb, TC_BK, BN);
    else rc = make_map(&tmB, g->B, g->N, g->K, g->ldb, 64, TC_BK);
    if (rc) return rc;
  } else {
    if (!A_MN) rc = make_map4(&tmA, g->A, g->K, g->M, g->lda, g->batch2, g->sA2, g->batch1, g->sA1, TC_BK, TC_BM);
    else rc = make_map4(&tmA, g->A, g->M, g->K, g->lda, g->batch2, g->sA2, g->batch1, g->sA1, 64, TC_BK);
    if (rc) return rc;
    if (!B_MN) rc = make_map4(&tmB, g->B, g->K, g->N, g->ldb, g->batch2, g->sB2, g->batch1, g->sB1, TC_BK, BN);
    else rc = make_map4(&tmB, g->B, g->N, g->K, g->ldb, g->batch2, g->sB2, g->batch1, g->sB1, 64, TC_BK);
    if (rc) return rc;
  }
  TcSched sc;
  sc.kb_total = cdiv(g->K, TC_BK);
  int split = g->split_k > 1 ? g->split_k : 1;
  if (split > sc.kb_total) split = sc.kb_total;
  sc.kb_per = cdiv(sc.kb_total, split);
  split = cdiv(sc.kb_total, sc.kb_per);  // no empty slices
  sc.splits = split;
  sc.tiles_n = cdiv(g->N, BN);
  sc.tiles_m = cdiv(g->M, TC_BM);
  sc.batch2 = g->batch2;
  sc.nbatch = g->batch1 * g->batch2;
  sc.total_tiles = (long)sc.tiles_n * sc.tiles_m * sc.nbatch * split;
  sc.sC1 = g->sC1;
  sc.sC2 = g->sC2;
  sc.sBias1 = g->sBias1;
  sc.sR1 = g->sR1;
  sc.sS1 = g->sS1;
  const bool atomic = (split > 1 || g->accumulate);
  sc.tma_store = (!atomic && getenv("CMX_GEMM_NO_TMA_STORE") == nullptr) ? 1 : 0;
  sc.trace = g_tc_trace;
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  if (sc.tma_store) {
    rc = make_map_c(&tmC, g, BATCHED);
    if (rc) return rc;
  }
  constexpr int STAGE_BYTES = (TC_BM + BN) * TC_BK * 2;
  constexpr uint32_t TC_CSTAGE_BYTES = tc_cstage_bytes(EW, CPS);
  int stages = (int)((tc_smem_budget(CPS) - 4096 - (sc.tma_store ? TC_CSTAGE_BYTES : 0)) / STAGE_BYTES);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  sc.stages = stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + (sc.tma_store ? TC_CSTAGE_BYTES : 0) + 2048 + 1024 + 16 * stages + 128;
  static thread_local PerDeviceOnce attr_once;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, BATCHED, EW, CPS>;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) CMX_FAIL((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_once.mark();
  }
  Epi e2 = epi;
  e2.atomic = atomic ? 1 : 0;
  const long slots = (long)num_sms() * CPS;
  const long grid = sc.total_tiles < slots ? sc.total_tiles : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc_threads(EW));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cmx_use_pdl() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, e2, sc);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("gemm_tc_kernel");
  return 0;
}

// Tile width.  Default policy "auto" (measured on B200 over the MiT-B2 shapes, graph-replayed launches,
// scripts/gemm_shapes_bench.py): the two-CTA-per-SM shapes only - BN = 128, or BN = 64 when N <= 64 or when 128-wide
// tiles would leave SMs empty (< 148 tiles: the small-token stages, where more and narrower tiles are 15-25 % faster).
// BN = 256 / 160 (whole-SM CTAs) are within +-3 % of BN = 128 standalone except for a few huge weight gradients and
// cost the co-residency with the other streams' kernels; they stay reachable for experiments and tests through
// CMX_GEMM_TILE_POLICY=pad (least padded columns, ties to the wider tile) and CMX_GEMM_BN_MAX.
static int pick_bn(const CmxGemm* g) {
  const long N = g->N;
  const char* pol = getenv("CMX_GEMM_TILE_POLICY");
  const char* emax = getenv("CMX_GEMM_BN_MAX");
  int bn_max = emax ? atoi(emax) : 256;
  if (bn_max < 64) bn_max = 64;
  if (pol && pol[0] == 'p') {
    const int cands_k[4] = {256, 160, 128, 64};
    const int cands_mn[3] = {256, 128, 64};   // MN-major B tiles are built from 64-wide chunks
    const int* c = g->trans_b ? cands_mn : cands_k;
    const int nc = g->trans_b ? 3 : 4;
    int best = c[nc - 1];
    long best_cost = -1;
    for (int i = 0; i < nc; i++) {
      if (c[i] > bn_max) continue;
      const long cost = (N + c[i] - 1) / c[i] * c[i];
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c[i]; }
    }
    return best;
  }
  if (N <= 64 || bn_max < 128) return 64;
  const long kb_total = (g->K + TC_BK - 1) / TC_BK;
  long split = g->split_k > 1 ? g->split_k : 1;
  if (split > kb_total) split = kb_total;
  const long tiles128 = ((g->M + TC_BM - 1) / TC_BM) * ((N + 127) / 128) * g->batch1 * g->batch2 * split;
  static long thr = -1;   // CMX_GEMM_BN64_BELOW: use 64-wide tiles while 128-wide ones would give fewer tiles than this
  if (thr < 0) {
    const char* e = getenv("CMX_GEMM_BN64_BELOW");
    thr = e ? atol(e) : num_sms();
  }
  return tiles128 < thr ? 64 : 128;
}

// epilogue warps of the two-CTA-per-SM shapes: 4 (default) or 8 (CMX_GEMM_EPI_WARPS=8).  Measured on B200 with the branch-free
// epilogue bodies (scripts/gpu_runs/r2_call9.sh): 8 warps need __launch_bounds__(320, 2) = 96 registers and spill 40-60 bytes;
// the step runs 20.80 ms with 4 warps against 21.1 ms with 8 (standalone the K <= 256 shapes are equal or faster with 4).
static int epi_warps_env() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("CMX_GEMM_EPI_WARPS");
    v = (s && atoi(s) == 8) ? 8 : 4;
  }
  return v;
}

// K from which the BN = 128 forward / data-gradient GEMMs take the deep single-CTA variant (CMX_GEMM_DEEP_K; default 0 = never).
// EXPERIMENT THAT DID NOT PAY (scripts/gpu_runs/r2_call17.sh, one B200): 5 stages in one CTA per SM are 10-25 % SLOWER than two
// co-resident CTAs with 2 stages each (M=153600 N=K=512: 134 vs 110 us; M=19200 N=320 K=1280: 44 vs 36 us; step 20.30 vs
// 20.02 ms) - the long-K shapes are bound by operand traffic out of L2, not by the depth of the ring.
static long deep_k_env() {
  static long v = -1;
  if (v < 0) {
    const char* s = getenv("CMX_GEMM_DEEP_K");
    v = s ? atol(s) : 0;
    if (v <= 0) v = 1L << 60;
  }
  return v;
}

static int dispatch_tc(const CmxGemm* g, const Epi& epi, cudaStream_t st) {
  const int bn = pick_bn(g);
  const bool a = g->trans_a != 0, b = g->trans_b != 0;
  const bool batched = g->batch1 != 1 || g->batch2 != 1;
  const bool atomic = g->split_k > 1 || g->accumulate;
  // deep variant: long K, plain (non split-K) output, and enough tiles that one CTA per SM still fills the machine
  const long tiles = ((g->M + TC_BM - 1) / TC_BM) * ((g->N + 127) / 128) * g->batch1 * g->batch2;
  const bool deep = bn == 128 && !a && !atomic && g->K >= deep_k_env() && tiles >= num_sms();
  const int ew = deep ? 8 : (bn <= 128 ? epi_warps_env() : 8);
  const int cps = (bn > 128 || deep) ? 1 : 2;
#define TC_CASE3(BNv, EWv, CPSv)                                                                       \
  if (bn == BNv && ew == EWv && cps == CPSv) {                                                         \
    if (!batched) {                                                                                    \
      if (!a && !b) return launch_tc<BNv, false, false, false, EWv, CPSv>(g, epi, st);                 \
      if (!a && b) return launch_tc<BNv, false, true, false, EWv, CPSv>(g, epi, st);                   \
      if constexpr (CPSv == 2 || BNv > 128) return launch_tc<BNv, true, true, false, EWv, CPSv>(g, epi, st); \
    } else {                                                                                           \
      if (!a && !b) return launch_tc<BNv, false, false, true, EWv, CPSv>(g, epi, st);                  \
      if (!a && b) return launch_tc<BNv, false, true, true, EWv, CPSv>(g, epi, st);                    \
      if constexpr (CPSv == 2 || BNv > 128) return launch_tc<BNv, true, true, true, EWv, CPSv>(g, epi, st); \
    }                                                                                                  \
  }
  TC_CASE3(64, 4, 2)
  TC_CASE3(128, 4, 2)
  TC_CASE3(128, 8, 1)
  TC_CASE3(64, 8, 2)
  TC_CASE3(128, 8, 2)
  TC_CASE3(256, 8, 1)
#undef TC_CASE3
  if (bn == 160) return batched ? launch_tc<160, false, false, true, 8, 1>(g, epi, st) : launch_tc<160, false, false, false, 8, 1>(g, epi, st);
  CMX_FAIL(-4, "no tcgen05 instantiation for BN=%d", bn);
}

static int force_impl_env() {
  static int v = -1;
  if (v < 0) {
    const char* s = getenv("CMX_GEMM_IMPL");
    v = 0;
    if (s && !strcmp(s, "fallback")) v = 1;
    if (s && !strcmp(s, "tc")) v = 2;
  }
  return v;
}

CMX_API int cmx_gemm_which(const CmxGemm* g) {
  int impl = g->impl ? g->impl : force_impl_env();
  if (impl == 1) return 1;
  return tc_eligible(g) ? 2 : 1;
}

CMX_API int cmx_gemm(const CmxGemm* g, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CMX_REQUIRE(g && g->A && g->B && g->C, "cmx_gemm: null operand");
  CMX_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "cmx_gemm: empty problem %ld x %ld x %ld", (long)g->M, (long)g->N, (long)g->K);
  CMX_REQUIRE(g->batch1 >= 1 && g->batch2 >= 1, "cmx_gemm: bad batch");
  CMX_REQUIRE(!((g->split_k > 1 || g->accumulate) && g->c_dtype != CMX_F32), "cmx_gemm: split-K/accumulate needs fp32 C");
  CMX_REQUIRE(!(g->row_scale && g->rows_per_sample <= 0), "cmx_gemm: row_scale needs rows_per_sample");
  CMX_REQUIRE(g->act == CMX_ACT_NONE || g->act == CMX_ACT_RELU, "cmx_gemm: epilogue activation must be none (0) or ReLU (1)");
  CMX_REQUIRE(g->M < (1ll << 31) && g->N < (1ll << 31), "cmx_gemm: M, N must fit in 31 bits");
  Epi epi;
  epi.C = g->C; epi.ldc = g->ldc; epi.bias = g->bias; epi.res = g->residual; epi.ldr = g->ldr;
  epi.row_scale = g->row_scale; epi.rows_per_sample = g->rows_per_sample > 0 ? g->rows_per_sample : 1;
  epi.c_dtype = g->c_dtype; epi.r_dtype = g->r_dtype; epi.act = g->act;
  epi.atomic = (g->split_k > 1 || g->accumulate) ? 1 : 0;
  epi.alpha = g->alpha; epi.M = g->M; epi.N = g->N;

  int impl = g->impl ? g->impl : force_impl_env();
  const bool elig = tc_eligible(g);
  if (impl == 2 && !elig) CMX_FAIL(-5, "cmx_gemm: tcgen05 path required but problem not eligible");
  if (impl != 1 && elig) return dispatch_tc(g, epi, st);

  GenArgs a;
  a.A = reinterpret_cast<const bf16*>(g->A);
  a.B = reinterpret_cast<const bf16*>(g->B);
  a.sAm = g->trans_a ? 1 : g->lda; a.sAk = g->trans_a ? g->lda : 1;
  a.sBk = g->trans_b ? g->ldb : 1; a.sBn = g->trans_b ? 1 : g->ldb;
  a.K = g->K; a.batch2 = g->batch2;
  a.sA1 = g->sA1; a.sA2 = g->sA2; a.sB1 = g->sB1; a.sB2 = g->sB2; a.sC1 = g->sC1; a.sC2 = g->sC2;
  a.sBias1 = g->sBias1; a.sR1 = g->sR1; a.sS1 = g->sS1;
  int split = g->split_k > 1 ? g->split_k : 1;
  long kper = ((g->K + split - 1) / split + GK - 1) / GK * GK;
  a.k_per_split = kper;
  split = (int)((g->K + kper - 1) / kper);
  epi.atomic = (split > 1 || g->accumulate) ? 1 : 0;
  dim3 grid(cdiv(g->N, GB), cdiv(g->M, GB), (unsigned)(g->batch1 * g->batch2 * split));
  CMX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "cmx_gemm fallback: grid too large");
  gemm_wmma_kernel<<<grid, 128, 0, st>>>(a, epi);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("gemm_wmma_kernel");
  return 0;
}
