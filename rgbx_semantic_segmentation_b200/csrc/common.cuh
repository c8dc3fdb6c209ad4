// Shared helpers for the cmx_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <atomic>

// cudaFuncSetAttribute is a per-DEVICE setting: a call site keeps one static PerDeviceOnce and repeats the opt-in for every
// device ordinal it is first used on (a second GPU in the same process would otherwise launch without its shared-memory opt-in).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};
  unsigned long long bit = 0;
  bool pending() {
    int d = 0;
    cudaGetDevice(&d);
    bit = 1ull << (d & 63);
    return !(done.load(std::memory_order_acquire) & bit);
  }
  void mark() { done.fetch_or(bit, std::memory_order_release); }
};

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local message, returned through cmx_last_error()) -------------
void cmx_set_error(const char* fmt, ...);
#define CMX_FAIL(code, ...)        \
  do {                             \
    cmx_set_error(__VA_ARGS__);    \
    return (code);                 \
  } while (0)
#define CMX_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) CMX_FAIL(-1, __VA_ARGS__);    \
  } while (0)
#define CMX_CHECK_LAUNCH(name)                                                    \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) CMX_FAIL((int)e__, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define CMX_API extern "C" __attribute__((visibility("default")))

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- dtype tags used across the C ABI ------------------------------------------------------
enum { CMX_BF16 = 0, CMX_F32 = 1 };
enum { CMX_ACT_NONE = 0, CMX_ACT_RELU = 1, CMX_ACT_GELU = 2 };

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// Every kernel signals "dependents may be scheduled" as its first instruction; only kernels launched with the
// programmatic-stream-serialization attribute (the tcgen05 GEMM and attention kernels) react to it: they run their
// prologue (barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel drains and then block in
// pdl_wait() until that kernel has fully completed and flushed its writes.  For ordinary launches both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- division of a 31-bit index by a runtime constant: q = umulhi(n, mul) >> shr (two instructions instead of the ~20 of a
// 32-bit hardware-less division; the movers decode a linear thread index into 4-6 coordinates per 16 bytes moved) -----------------
struct FastDiv {
  unsigned d, mul, shr;
};
static inline FastDiv make_fastdiv(unsigned d) {   // valid for dividends < 2^31
  FastDiv f;
  f.d = d;
  f.mul = 0;
  f.shr = 0;
  if (d > 1) {
    unsigned lg = 0;
    while ((1ull << lg) < d) lg++;
    const unsigned p = 31 + lg;
    f.mul = (unsigned)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}
__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv& f) { return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr); }
__device__ __forceinline__ void fdivmod(unsigned n, const FastDiv& f, unsigned& q, unsigned& r) {
  q = fdiv(n, f);
  r = n - q * f.d;
}

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU (nn.GELU default, dual_segformer.py:21) evaluated with the Abramowitz-Stegun 7.1.26 rational
// approximation of erf: |erf error| <= 1.5e-7, i.e. |gelu error| <= 4.3e-7 and |gelu' error| <= 3.2e-7 over [-12, 12]
// (checked against scipy in float32) - far below the bf16 resolution of the tensors it is applied to - in ~16
// instructions (one MUFU.EX2, one MUFU.RCP) instead of erff's ~35.  exp(-z^2) with z = |x|/sqrt(2) is also the
// Gaussian pdf factor, so the derivative shares it.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& e) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  e = __expf(-z * z);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float h = 0.5f * p * t * e;  // 0.5 * erfc(z)
  cdf = x >= 0.f ? 1.f - h : h;
}
__device__ __forceinline__ float gelu_f(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}
// packed fp32x2 helpers (Blackwell FFMA2 / FMUL2 / FADD2: both halves in ONE issue slot)
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(const float2 a, const float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 ffma2r(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// GELU / GELU' of a channel pair with the packed ops (same A&S 7.1.26 erf as gelu_parts): ~19 / ~21 issue slots per
// PAIR instead of 2 x 16 / 2 x 18
__device__ __forceinline__ void gelu_parts2(const float2 x, float2& cdf, float2& e) {
  const float2 x2 = fmul2(x, x);
  const float2 arg = fmul2(x2, make_float2(-0.72134752044448170368f, -0.72134752044448170368f));  // -0.5 * log2(e) * x^2
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));  // ONE MUFU.EX2 each (exp2f() adds a range fix-up: 3 more
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));  // instructions; arg <= 0 here and results below 2^-126 are 0 anyway)
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x.x), 1.f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x.y), 1.f)));
  // 0.5 * (a1 t + ... + a5 t^5) = ((((h5 t + h4) t + h3) t + h2) t + h1) t with h_i = a_i / 2
  float2 p = ffma2r(make_float2(0.5307027145f, 0.5307027145f), t, make_float2(-0.7265760135f, -0.7265760135f));
  p = ffma2r(p, t, make_float2(0.7107068705f, 0.7107068705f));
  p = ffma2r(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = ffma2r(p, t, make_float2(0.127414796f, 0.127414796f));
  const float2 h = fmul2(fmul2(p, t), e);  // 0.5 * erfc(|x| / sqrt 2)
  // cdf = 0.5 + copysign(0.5 - h, x)
  float2 s = fadd2(make_float2(0.5f, 0.5f), make_float2(-h.x, -h.y));
  s.x = copysignf(s.x, x.x);
  s.y = copysignf(s.y, x.y);
  cdf = fadd2(make_float2(0.5f, 0.5f), s);
}
__device__ __forceinline__ float2 gelu2(const float2 x) {
  float2 cdf, e;
  gelu_parts2(x, cdf, e);
  return fmul2(x, cdf);
}
__device__ __forceinline__ float2 gelu_grad2(const float2 x) {
  float2 cdf, e;
  gelu_parts2(x, cdf, e);
  return ffma2r(fmul2(x, make_float2(0.39894228040143267794f, 0.39894228040143267794f)), e, cdf);
}
// packed fp32x2 FMA: d = a * b + d
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rd, {%0, %1};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rd;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d.x), "+f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// 8 x bf16 <-> 8 x float through one 16-byte access
struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const bf16* p, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void load4(const bf16* p, float* f) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void load4(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
__device__ __forceinline__ void store4(bf16* p, const float* f) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(f[0], f[1]);
  h[1] = __floats2bfloat162_rn(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store4(float* p, const float* f) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(bf16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

// PyTorch's bilinear source index (align_corners=False): aten/src/ATen/native/UpSample.h
// area_pixel_compute_source_index.  scale = in/out in fp32.
__device__ __forceinline__ void bilin_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
}
