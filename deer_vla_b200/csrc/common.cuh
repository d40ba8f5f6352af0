// Common device helpers for the DeeR B200 engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

typedef __nv_bfloat16 bf16;

#define DEER_NUM_SMS 148

// ---- scalar conversions ---------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- activations (exact forms used by the reference) ------------------------------------
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_QUICKGELU = 2, ACT_RELU = 3 };

__device__ __forceinline__ float gelu_erf(float x) {          // nn.GELU() (approximate='none')
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float quick_gelu(float x) {        // open_clip QuickGELU: x*sigmoid(1.702x)
  return x / (1.0f + __expf(-1.702f * x));
}
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_QUICKGELU) return quick_gelu(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}
// Fast forms for the tensor-core epilogue (results are rounded to bf16 or added to an O(1) fp32 residual).
// one MUFU.EX2, no range fix-up
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact (erf) GELU with ONE MUFU op and no reciprocal: gelu(x) = x * Phi(x) with Phi(-a) = 0.5 * erfc(a / sqrt 2) = 2^L(a).
// L is a degree-7 polynomial fitted on [0, 6] (max |dL| 5.7e-6, L(0) = -1 exactly; beyond 6, Phi(-a) < 1e-9 is clamped):
//   x < 0: x * 2^L(|x|)        x >= 0: x - x * 2^L(x)
// so the negative tail keeps RELATIVE accuracy (no 1 - erf cancellation).  |err| <= 7e-7 absolute in fp32 over [-8, 8]
// (fit + error check: DESIGN.md section 4); the previous Abramowitz-Stegun 7.1.26 form needed a reciprocal and an exp.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float p = -1.793804472e-06f;
  p = fmaf(p, a, 6.072663382e-05f);
  p = fmaf(p, a, -9.231893346e-04f);
  p = fmaf(p, a, 8.476860821e-03f);
  p = fmaf(p, a, -5.389096215e-02f);
  p = fmaf(p, a, -4.585415721e-01f);
  p = fmaf(p, a, -1.151214361e+00f);
  p = fmaf(p, a, -1.0f);
  const float xt = x * ex2_approx(p);
  return x >= 0.0f ? x - xt : xt;
}
// x*sigmoid(1.702x) with sigmoid(y) = 0.5 + 0.5*tanh(y/2): ONE MUFU op per element (tanh.approx.f32, rel. err 2^-11,
// i.e. <= 2.5e-4*|x| absolute -- an order of magnitude below the bf16 rounding of the stored value)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quick_gelu_fast(float x) { return x * fmaf(0.5f, tanh_approx(0.851f * x), 0.5f); }
template <int ACT> __device__ __forceinline__ float act_fast(float v) {
  if (ACT == ACT_GELU) return gelu_erf_fast(v);
  if (ACT == ACT_QUICKGELU) return quick_gelu_fast(v);
  if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in the stream is
// still draining: its prologue (barrier init, TMEM allocation, tensormap prefetch) overlaps the predecessor's tail.  pdl_wait()
// blocks until the predecessor grid has completed and its writes are visible -- it must precede the first global access that
// depends on earlier work (including the device-side row count); without the launch attribute it returns immediately.
// pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as this grid's CTAs free their SM resources.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- warp / block reductions ---------------------------------------------------------------
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

// dynamic row count: rows = min(M_host, *m_dev * m_mul) when m_dev != nullptr
__device__ __forceinline__ int dyn_rows(int m_host, const int* m_dev, int m_mul) {
  if (m_dev == nullptr) return m_host;
  int m = (*m_dev) * m_mul;
  return m < m_host ? m : m_host;
}
