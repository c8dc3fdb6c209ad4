// AdamW over the engine's flat parameter storage: ONE launch for all 522+ parameters of the model (the reference
// steps torch.optim.AdamW over parameter groups, train.py:95-100; torch's fused multi-tensor path needs ~30 launches
// with partially filled grids for the same update).  Memory-bound: 16 B read + 12 B written per element.
#include "common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>
extern std::atomic<long long> g_cmx_launches;

constexpr int AW_MAXG = 8;
struct AdamWArgs {
  float lr[AW_MAXG], wd[AW_MAXG];
  float beta1, beta2, eps, bc1, bc2_rsqrt, grad_scale;
};

// one thread = 4 consecutive elements (16-byte accesses); the hyper-parameter group is a property of the 64-element
// block (the flat layout aligns every parameter to 64 elements), group 255 = not optimised (frozen / padding)
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, bf16* __restrict__ w,
                                                         const uint8_t* __restrict__ block_group, long n4, AdamWArgs a) {
  pdl_trigger();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const int grp = block_group[i >> 4];
    if (grp == 255) continue;
    const float lr = a.lr[grp], wd = a.wd[grp];
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x;
    const float* gp = &gv.x;
    float* mp = &mv.x;
    float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float gr = gp[k] * a.grad_scale;
      // torch/optim/adamw.py (single-tensor path): p *= 1 - lr*wd; m, v moments; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
      float x = pp[k] * (1.f - lr * wd);
      mp[k] = fmaf(a.beta1, mp[k], (1.f - a.beta1) * gr);
      vp[k] = fmaf(a.beta2, vp[k], (1.f - a.beta2) * gr * gr);
      const float denom = fmaf(sqrtf(vp[k]), a.bc2_rsqrt, a.eps);
      x -= (lr / a.bc1) * (mp[k] / denom);
      pp[k] = x;
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (w) {
      float o[4] = {pv.x, pv.y, pv.z, pv.w};
      store4(w + 4 * i, o);
    }
  }
}

CMX_API int cmx_adamw_flat(float* p, const float* g, float* m, float* v, void* w_bf16, const uint8_t* block_group, int64_t n,
                           const float* lr, const float* wd, int ngroups, float beta1, float beta2, float eps, float grad_scale,
                           int64_t step, void* stream) {
  CMX_REQUIRE(n % 64 == 0, "adamw_flat: n=%ld must be a multiple of the 64-element block", (long)n);
  CMX_REQUIRE(ngroups >= 1 && ngroups <= AW_MAXG, "adamw_flat: %d parameter groups (max %d)", ngroups, AW_MAXG);
  CMX_REQUIRE(step >= 1, "adamw_flat: step counts from 1");
  CMX_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0 && ((uintptr_t)w_bf16 & 7) == 0,
              "adamw_flat: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  AdamWArgs a;
  for (int i = 0; i < AW_MAXG; i++) { a.lr[i] = i < ngroups ? lr[i] : 0.f; a.wd[i] = i < ngroups ? wd[i] : 0.f; }
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_rsqrt = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
  const long n4 = n / 4;
  long grid = (n4 + 255) / 256;
  if (grid > 148L * 16) grid = 148L * 16;
  adamw_flat_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)w_bf16, block_group, n4, a);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("adamw_flat");
  return 0;
}
