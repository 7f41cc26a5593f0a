// libwctb: default fp32-product Gram kernel for C = 24 / 32 -- register accumulation fed from a cp.async ring.
#include "gram_small.cuh"

// ------------------------------------------------------------------------------------------
// Same register-resident accumulation, fed from a cp.async ring in shared memory (variant 0, default).
// gram_regs_kernel re-reads every pixel once per warp group through L1 and consumes each load immediately, which leaves
// it latency-bound at a third of the HBM rate.  Here the CTA copies each pixel tile ([NCH][PIX] float4) exactly once
// with cp.async.cg (16 B per thread and copy, coalesced) into a RING-deep ring; all groups read their pixel's chunks from
// it with conflict-free LDS.128 (lane j <-> pixel j).  One barrier per iteration: wait for tile `it`, barrier, refill the
// slot that was computed in iteration it-1, compute tile `it`.  RING-1 tiles (~45 KB) are in flight per SM.
// ------------------------------------------------------------------------------------------
// copy pixel tile `t` (pixels blockIdx.x*PIX + t*stride + [0,PIX)) into ring slot t % RING; pixels beyond npix are skipped
// (their slots are never used: the compute step range-checks the last tile).  Executed by all PIX*SPLIT threads.
template <int NCH, int SPLIT, int PIX, bool FULLROW>
__device__ __forceinline__ void gram_ring_issue(float4* __restrict__ ring_slot, const float4* __restrict__ x, long long HW,
                                                int W, int y0, int x0, unsigned wreg, unsigned npix, unsigned tile_base) {
  constexpr int NT = PIX * SPLIT;
#pragma unroll
  for (int e0 = 0; e0 < NCH * PIX; e0 += NT) {
    const int e = e0 + (int)threadIdx.x;
    if (e < NCH * PIX) {
      const int c = e / PIX, j = e - c * PIX;
      const unsigned p = tile_base + j;
      if (p < npix) {
        long long off;
        if (FULLROW) {
          off = (long long)y0 * W + p;
        } else {
          const unsigned r = p / wreg, cc = p - r * wreg;
          off = (long long)(y0 + r) * W + (x0 + cc);
        }
        cp_async16(ring_slot + e, x + (long long)c * HW + off);
      }
    }
  }
}

template <int NCH, int SPLIT, int PART, int PIX, int RING, bool FULLROW, bool PEEL>
__device__ __forceinline__ void gram_ring_body(float4* __restrict__ ring, const float4* __restrict__ x, long long HW, int W, int y0,
                                               int x0, unsigned wreg, unsigned npix, unsigned iters,
                                               const float4* __restrict__ s_mh, double* __restrict__ G) {
  using D = GramDeal<NCH, SPLIT>;
  constexpr int C = NCH * 4;
  constexpr int NP = D::MAXCOUNT;
  constexpr int TILE = NCH * PIX;   // float4 per ring slot
  float acc[NP][16];
#pragma unroll
  for (int s = 0; s < NP; ++s)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[s][e] = 0.f;
  const unsigned stride = gridDim.x * PIX;
  const unsigned cta_base = blockIdx.x * PIX;
  const int j = threadIdx.x % PIX;
  // prologue: tiles 0 .. RING-2 (one commit group per tile, empty groups keep the count uniform)
#pragma unroll
  for (int t = 0; t < RING - 1; ++t) {
    if ((unsigned)t < iters) gram_ring_issue<NCH, SPLIT, PIX, FULLROW>(ring + t * TILE, x, HW, W, y0, x0, wreg, npix, cta_base + t * stride);
    cp_async_commit();
  }
  // PEEL (variant 3, not yet the default): the range-checked last tile is handled after the loop, which halves the loop's
  // code (ncu: 0.59 warps stalled on instruction fetch per issue in the unpeeled loop, profiles/r01_gram_ring_ncu_full.txt)
  const unsigned loop_iters = PEEL ? iters - 1 : iters;
  for (unsigned it = 0; it < loop_iters; ++it) {
    cp_async_wait<RING - 2>();       // this thread's copies of tile `it` have landed
    // CTA-wide named barrier (the groups sit in different branches of the dispatch): everybody's copies of tile `it` are
    // visible, and everybody has finished computing tile it-1, whose slot is refilled next
    asm volatile("bar.sync 1, %0;" ::"r"(PIX * SPLIT) : "memory");
    const unsigned tn = it + (RING - 1);
    if (tn < iters) gram_ring_issue<NCH, SPLIT, PIX, FULLROW>(ring + (tn % RING) * TILE, x, HW, W, y0, x0, wreg, npix, cta_base + tn * stride);
    cp_async_commit();
    const float4* tile = ring + (it % RING) * TILE;
    if (PEEL || it + 1 < iters) {
      gram_ring_step<NCH, SPLIT, PART, PIX, false>(acc, tile, j, true, s_mh);
    } else {
      gram_ring_step<NCH, SPLIT, PART, PIX, true>(acc, tile, j, cta_base + it * stride + j < npix, s_mh);
    }
  }
  if (PEEL) {
    const unsigned it = iters - 1;
    cp_async_wait<0>();
    asm volatile("bar.sync 1, %0;" ::"r"(PIX * SPLIT) : "memory");
    gram_ring_step<NCH, SPLIT, PART, PIX, true>(acc, ring + (it % RING) * TILE, j, cta_base + it * stride + j < npix, s_mh);
  }
  cp_async_wait<0>();
  // flush
  const int lane = threadIdx.x & 31;
  {
    int q = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
#pragma unroll
      for (int jj = i; jj < NCH; ++jj) {
        if (D::owns(PART, q)) {
          const int slot = q - D::begin(PART);
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              double sum = (double)acc[slot][u * 4 + v];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              if (lane == 0) {
                atomicAdd(G + (long long)(i * 4 + u) * C + (jj * 4 + v), sum);
                if (i != jj) atomicAdd(G + (long long)(jj * 4 + v) * C + (i * 4 + u), sum);
              }
            }
        }
        ++q;
      }
    }
  }
}

template <int NCH, int SPLIT, int PART, int PIX, int RING, bool FULLROW, bool PEEL>
__device__ __forceinline__ void gram_ring_dispatch(int part, float4* __restrict__ ring, const float4* __restrict__ x, long long HW,
                                                   int W, int y0, int x0, unsigned wreg, unsigned npix, unsigned iters,
                                                   const float4* s_mh, double* __restrict__ G) {
  if (part == PART) {
    gram_ring_body<NCH, SPLIT, PART, PIX, RING, FULLROW, PEEL>(ring, x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  } else if constexpr (PART + 1 < SPLIT) {
    gram_ring_dispatch<NCH, SPLIT, PART + 1, PIX, RING, FULLROW, PEEL>(part, ring, x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  }
}

template <int NCH, int SPLIT, int PIX, int RING, bool FULLROW, bool PEEL>
__global__ void __launch_bounds__(PIX* SPLIT, 1) gram_ring_kernel(const float4* __restrict__ x, int H, int W, int y0, int x0,
                                                                   unsigned wreg, unsigned npix, unsigned iters,
                                                                   const double* __restrict__ mean, double* __restrict__ G) {
  static_assert(PIX % 32 == 0, "a warp must not straddle two groups");
  static_assert(RING >= 3, "ring too shallow");
  extern __shared__ __align__(16) unsigned char ring_raw[];
  float4* ring = reinterpret_cast<float4*>(ring_raw);   // [RING][NCH][PIX]
  __shared__ float4 s_mh[NCH];
  if (threadIdx.x < NCH * 4) reinterpret_cast<float*>(s_mh)[threadIdx.x] = (float)mean[threadIdx.x];
  __syncthreads();
  gram_ring_dispatch<NCH, SPLIT, 0, PIX, RING, FULLROW, PEEL>(threadIdx.x / PIX, ring, x, (long long)H * W, W, y0, x0, wreg, npix, iters,
                                                        s_mh, G);
}

template <int NCH, int SPLIT, int PIX, int RING, bool PEEL>
static int launch_gram_ring(const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean, double* gram_out,
                            cudaStream_t st) {
  const long long npix = (long long)(y1 - y0) * (x1 - x0);
  long long ctas = (npix + PIX - 1) / PIX;
  const long long cap = wctb_num_sms();          // one persistent CTA per SM (register-limited)
  if (ctas > cap) ctas = cap;
  const long long stride = ctas * PIX;
  const unsigned iters = (unsigned)((npix + stride - 1) / stride);   // (iters-1)*stride < npix: only the last tile is partial
  const unsigned wreg = (unsigned)(x1 - x0);
  const size_t smem = (size_t)RING * NCH * PIX * sizeof(float4);
  static bool attr_done[WCTB_MAX_DEVICES] = {};      // per device: one process may drive several GPUs
  const int dev_slot = wctb_device_slot();
  if (!attr_done[dev_slot]) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(gram_ring_kernel<NCH, SPLIT, PIX, RING, true, PEEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WCTB_CUDA_TRY(cudaFuncSetAttribute(gram_ring_kernel<NCH, SPLIT, PIX, RING, false, PEEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev_slot] = true;
  }
  if (x0 == 0 && x1 == W)
    gram_ring_kernel<NCH, SPLIT, PIX, RING, true, PEEL><<<(unsigned)ctas, PIX * SPLIT, smem, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                                   (unsigned)npix, iters, mean, gram_out);
  else
    gram_ring_kernel<NCH, SPLIT, PIX, RING, false, PEEL><<<(unsigned)ctas, PIX * SPLIT, smem, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                                    (unsigned)npix, iters, mean, gram_out);
  WCTB_RETURN_LAUNCH();
}


int wctb_gram_ring_launch(int C, int peel, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                          double* gram_out, cudaStream_t st) {
  if (C == 24) return peel ? launch_gram_ring<6, 4, 96, 6, true>(x, H, W, y0, y1, x0, x1, mean, gram_out, st)
                           : launch_gram_ring<6, 4, 96, 6, false>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C == 32) return peel ? launch_gram_ring<8, 6, 64, 6, true>(x, H, W, y0, y1, x0, x1, mean, gram_out, st)
                           : launch_gram_ring<8, 6, 64, 6, false>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  return WCTB_E_UNSUPPORTED;
}
