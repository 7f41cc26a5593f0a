// Shared helpers for libwctb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wctb.h"

extern thread_local int g_wctb_last_cuda_error;

static inline int wctb_check_launch_() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_wctb_last_cuda_error = (int)e;
    return WCTB_E_CUDA;
  }
  return WCTB_OK;
}
#define WCTB_RETURN_LAUNCH() return wctb_check_launch_()
#define WCTB_CUDA_TRY(expr)                 \
  do {                                      \
    cudaError_t e__ = (expr);               \
    if (e__ != cudaSuccess) {               \
      g_wctb_last_cuda_error = (int)e__;    \
      return WCTB_E_CUDA;                   \
    }                                       \
  } while (0)

// nn.ReflectionPad2d(1): -1 -> 1, n -> n-2 (no edge duplication).  Also clamps far
// out-of-range indices (only reached by masked-out lanes of partial tiles).
__host__ __device__ __forceinline__ int wctb_reflect(int i, int n) {
  i = i < 0 ? -i : i;
  i = i >= n ? 2 * n - 2 - i : i;
  i = i < 0 ? 0 : i;
  i = i >= n ? n - 1 : i;
  return i;
}

// round-to-nearest (ties away) to TF32, kept in an fp32 container
__device__ __forceinline__ float wctb_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float wctb_relu(float v) { return v > 0.f ? v : 0.f; }

#define WCTB_MAX_DEVICES 64
static inline int wctb_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < WCTB_MAX_DEVICES) ? dev : 0;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel instantiation, device); `flags` is a static
// bool[WCTB_MAX_DEVICES] owned by the launcher of that instantiation (immutable kernel attributes are the only state)
#define WCTB_SET_SMEM_ONCE(flags, kernel, bytes)                                                              \
  do {                                                                                                        \
    const int slot__ = wctb_device_slot();                                                                    \
    if (!(flags)[slot__]) {                                                                                   \
      WCTB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      (flags)[slot__] = true;                                                                                 \
    }                                                                                                         \
  } while (0)

static inline int wctb_num_sms() {
  static int n[WCTB_MAX_DEVICES] = {};
  const int slot = wctb_device_slot();
  if (n[slot] == 0) {
    cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, slot);
    if (n[slot] <= 0) n[slot] = 148;
  }
  return n[slot];
}
