// Shared device helpers of the h2 (fp16 hi/lo pair) convolution engine: conv_h2.cu, conv_h2_fused.cu.
#pragma once
#include <cuda_fp16.h>

#include "umma.cuh"

namespace wctb_umma {

__device__ __forceinline__ uint32_t f2h_sat(float v) {   // fp32 -> fp16 bits, round-to-nearest-even, saturating
  uint16_t h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return h;
}
__device__ __forceinline__ float h2f(uint32_t h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((uint16_t)h));
  return f;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
// packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2 -- one issue slot for two fp32 operations; the epilogues of the fused
// kernels are bound by instruction issue, not by the tensor pipe)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// two floats -> packed fp16x2 (a in the low half, b in the high half), round-to-nearest-even, saturating: one F2FP
__device__ __forceinline__ uint32_t f2h2_sat(float a, float b) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ void h22f(uint32_t h2, float& a, float& b) {
  const __half2 h = *reinterpret_cast<const __half2*>(&h2);
  const float2 f = __half22float2(h);
  a = f.x;
  b = f.y;
}
// x -> (hi, lo) fp16 pair; 8 values -> two 16-byte units
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = f2h2_sat(v[2 * i], v[2 * i + 1]);
    float a, b;
    h22f(h[i], a, b);
    l[i] = f2h2_sat(v[2 * i] - a, v[2 * i + 1] - b);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// (hi, lo) 16-byte units -> 8 floats
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float* v) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int e = 0; e < 8; ++e)
    v[e] = h2f((h[e >> 1] >> (16 * (e & 1))) & 0xffffu) + h2f((l[e >> 1] >> (16 * (e & 1))) & 0xffffu);
}

// cudaFuncSetAttribute once per (kernel, device)
template <class K>
static int ensure_smem_attr(K kernel, int bytes, bool* done) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!done[dev]) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done[dev] = true;
  }
  return WCTB_OK;
}


}  // namespace wctb_umma
