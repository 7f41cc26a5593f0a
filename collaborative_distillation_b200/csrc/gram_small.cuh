// Shared pieces of the register-resident fp32 Gram kernels (gram_ring.cu, gram_alt.cu); see wct_transform.cu for the
// staged kernels and the dispatch (wctb_centered_gram_fast).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// Register-resident fp32 Gram for the small-channel, many-pixel stages (C = 24: stage 1, C = 32: stage 2).
// The staged kernel above is shared-memory bound there (6 LDS.32 per 9 FFMA: CUDA events put it at ~5x its FFMA and
// HBM floors).  Here a thread owns whole PIXELS: it loads the chunks (float4 = 4 channels) of its pixel straight from
// global memory (a warp reads 512 contiguous bytes per chunk plane), centres them with a split (hi, lo) fp32 mean and
// accumulates 4x4 outer-product blocks of the upper block triangle in registers.  The NCH(NCH+1)/2 blocks are dealt in
// contiguous runs to SPLIT warp groups ("parts") of the CTA; a part only loads the chunks its blocks touch, all parts
// walk the same pixels (re-reads hit L1; a barrier every 4th iteration keeps the parts inside the L1 window), and each
// part prefetches its lines two iterations ahead.  No shared-memory traffic in the loop.  Register budget: 12 warps/SM
// (3 per scheduler) -> <= 168 registers: <= 96 accumulators + <= 32 operands.
// Flush: warp-shuffle reduction in fp64, then fp64 atomics (same accumulate-into-G contract as the staged kernel).
// Accuracy: per-thread fp32 sums over npix/(gridDim*PIX) pixels (hundreds..2e3) -> ~1e-6 each, averaging over the
// >= 1e4 threads to ~1e-8 relative in G; the contract of the fast variant is 1e-6 (tests/test_gpu_parity.py).
// ------------------------------------------------------------------------------------------
template <int NCH, int SPLIT>
struct GramDeal {
  static constexpr int NPAIR = NCH * (NCH + 1) / 2;
  static constexpr int BASE = NPAIR / SPLIT, REM = NPAIR % SPLIT;
  __host__ __device__ static constexpr int begin(int part) { return part * BASE + (part < REM ? part : REM); }
  __host__ __device__ static constexpr int count(int part) { return BASE + (part < REM ? 1 : 0); }
  static constexpr int MAXCOUNT = BASE + (REM ? 1 : 0);
  __host__ __device__ static constexpr bool owns(int part, int q) { return q >= begin(part) && q < begin(part) + count(part); }
  __host__ __device__ static constexpr bool uses_chunk(int part, int c) {
    int q = 0;
    for (int i = 0; i < NCH; ++i)
      for (int j = i; j < NCH; ++j, ++q)
        if (owns(part, q) && (i == c || j == c)) return true;
    return false;
  }
};


__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


template <int NCH, int SPLIT, int PART, int PIX, bool CHECK>
__device__ __forceinline__ void gram_ring_step(float (&acc)[GramDeal<NCH, SPLIT>::MAXCOUNT][16], const float4* __restrict__ tile,
                                               int j, bool have, const float4* __restrict__ s_mh) {
  using D = GramDeal<NCH, SPLIT>;
  float4 cur[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    cur[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (D::uses_chunk(PART, c) && (!CHECK || have)) {
      const float4 v = tile[c * PIX + j];
      const float4 mh = s_mh[c];
      cur[c] = make_float4(v.x - mh.x, v.y - mh.y, v.z - mh.z, v.w - mh.w);
    }
  }
  int q = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
#pragma unroll
    for (int jj = i; jj < NCH; ++jj) {
      if (D::owns(PART, q)) {
        const int slot = q - D::begin(PART);
        const float a[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
        const float b[4] = {cur[jj].x, cur[jj].y, cur[jj].z, cur[jj].w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[slot][u * 4 + v] = fmaf(a[u], b[v], acc[slot][u * 4 + v]);
      }
      ++q;
    }
  }
}


// launchers of the small-channel fp32 Gram kernels (C = 24 or 32; any other C returns WCTB_E_UNSUPPORTED).  Defined in
// gram_ring.cu (default: cp.async ring feed; peel != 0: last iteration peeled) and gram_alt.cu (L1 feed; two pixels per
// thread) -- separate translation units so that they compile side by side with wct_transform.cu.
int wctb_gram_ring_launch(int C, int peel, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                          double* gram_out, cudaStream_t st);
int wctb_gram_regs_launch(int C, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                          double* gram_out, cudaStream_t st);
int wctb_gram_ring2_launch(int C, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                           double* gram_out, cudaStream_t st);
