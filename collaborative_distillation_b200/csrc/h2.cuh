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
