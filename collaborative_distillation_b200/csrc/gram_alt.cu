// libwctb: alternative fp32-product Gram kernels for C = 24 / 32 kept for A/B timing (wctb_debug_set_gram_variant):
// variant 2 = register accumulation fed through L1, variant 4 = ring feed with two pixels per thread and iteration.
#include "gram_small.cuh"

// one pixel per thread: load the chunks this part touches, centre, accumulate its blocks.  CHECK = last iteration (slots
// beyond npix contribute zero); FULLROW = the region spans whole rows, so pixel p of the region is pixel y0*W + p of the plane.
template <int NCH, int SPLIT, int PART, bool FULLROW, bool CHECK>
__device__ __forceinline__ void gram_regs_step(float (&acc)[GramDeal<NCH, SPLIT>::MAXCOUNT][16], const float4* __restrict__ x,
                                               long long HW, int W, int y0, int x0, unsigned wreg, unsigned npix, unsigned p,
                                               const float4* __restrict__ s_mh) {
  using D = GramDeal<NCH, SPLIT>;
  float4 cur[NCH];
  const bool have = !CHECK || p < npix;
  long long off;
  if (FULLROW) {
    off = (long long)y0 * W + p;
  } else {
    const unsigned pp = have ? p : 0u;
    const unsigned r = pp / wreg, cc = pp - r * wreg;
    off = (long long)(y0 + r) * W + (x0 + cc);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    cur[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (D::uses_chunk(PART, c) && have) {
      const float4 v = __ldg(x + (long long)c * HW + off);
      const float4 mh = s_mh[c];
      // fp32 mean: its rounding error d (<= 2^-24 |mean|) only adds N d_i d_j to G, ~1e-14 relative
      cur[c] = make_float4(v.x - mh.x, v.y - mh.y, v.z - mh.z, v.w - mh.w);
    }
  }
  int q = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
#pragma unroll
    for (int j = i; j < NCH; ++j) {
      if (D::owns(PART, q)) {
        const int slot = q - D::begin(PART);
        const float a[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
        const float b[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[slot][u * 4 + v] = fmaf(a[u], b[v], acc[slot][u * 4 + v]);
      }
      ++q;
    }
  }
}

template <int NCH, int SPLIT, int PART, int PIX, bool FULLROW>
__device__ __forceinline__ void gram_regs_body(const float4* __restrict__ x, long long HW, int W, int y0, int x0, unsigned wreg,
                                               unsigned npix, unsigned iters, const float4* __restrict__ s_mh,
                                               double* __restrict__ G) {
  using D = GramDeal<NCH, SPLIT>;
  constexpr int C = NCH * 4;
  constexpr int NP = D::MAXCOUNT;
  float acc[NP][16];
#pragma unroll
  for (int s = 0; s < NP; ++s)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[s][e] = 0.f;
  const unsigned stride = gridDim.x * PIX;
  const unsigned pbase = blockIdx.x * PIX + (threadIdx.x % PIX);
  // iterations 0 .. iters-2 are in range for every thread (see launch_gram_regs); only the last one needs the range check.
  // Uniform trip count: the barrier is reached by every thread of the CTA.
  for (unsigned it = 0; it + 1 < iters; ++it) {
    const unsigned p = pbase + it * stride;          // no 32-bit overflow: the host checks npix < 2^31 - 2^24
    {   // prefetch two iterations ahead
      const unsigned pf = p + 2 * stride;
      if (pf < npix) {
        long long off;
        if (FULLROW) {
          off = (long long)y0 * W + pf;
        } else {
          const unsigned r = pf / wreg, cc = pf - r * wreg;
          off = (long long)(y0 + r) * W + (x0 + cc);
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (D::uses_chunk(PART, c)) asm volatile("prefetch.global.L1 [%0];" ::"l"(x + (long long)c * HW + off));
      }
    }
    gram_regs_step<NCH, SPLIT, PART, FULLROW, false>(acc, x, HW, W, y0, x0, wreg, npix, p, s_mh);
    // named barrier over the whole CTA: the parts sit in different branches of the dispatch (warp-uniform), so this is
    // written as bar.sync <id>, <count> rather than __syncthreads()
    if ((it & 3) == 3) asm volatile("bar.sync 1, %0;" ::"r"(PIX * SPLIT) : "memory");
  }
  gram_regs_step<NCH, SPLIT, PART, FULLROW, true>(acc, x, HW, W, y0, x0, wreg, npix, pbase + (iters - 1) * stride, s_mh);
  // flush
  const int lane = threadIdx.x & 31;
  {
    int q = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
#pragma unroll
      for (int j = i; j < NCH; ++j) {
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
                atomicAdd(G + (long long)(i * 4 + u) * C + (j * 4 + v), sum);
                if (i != j) atomicAdd(G + (long long)(j * 4 + v) * C + (i * 4 + u), sum);
              }
            }
        }
        ++q;
      }
    }
  }
}

template <int NCH, int SPLIT, int PART, int PIX, bool FULLROW>
__device__ __forceinline__ void gram_regs_dispatch(int part, const float4* __restrict__ x, long long HW, int W, int y0, int x0,
                                                   unsigned wreg, unsigned npix, unsigned iters, const float4* s_mh,
                                                   double* __restrict__ G) {
  if (part == PART) {
    gram_regs_body<NCH, SPLIT, PART, PIX, FULLROW>(x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  } else if constexpr (PART + 1 < SPLIT) {
    gram_regs_dispatch<NCH, SPLIT, PART + 1, PIX, FULLROW>(part, x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  }
}

template <int NCH, int SPLIT, int PIX, bool FULLROW>
__global__ void __launch_bounds__(PIX* SPLIT, 1) gram_regs_kernel(const float4* __restrict__ x, int H, int W, int y0, int x0,
                                                                   unsigned wreg, unsigned npix, unsigned iters,
                                                                   const double* __restrict__ mean, double* __restrict__ G) {
  static_assert(PIX % 32 == 0, "a warp must not straddle two parts");
  __shared__ float4 s_mh[NCH];
  if (threadIdx.x < NCH * 4) reinterpret_cast<float*>(s_mh)[threadIdx.x] = (float)mean[threadIdx.x];
  __syncthreads();
  // every part runs the same number of barriers (iters is uniform), so the divergent dispatch is barrier-safe
  gram_regs_dispatch<NCH, SPLIT, 0, PIX, FULLROW>(threadIdx.x / PIX, x, (long long)H * W, W, y0, x0, wreg, npix, iters, s_mh, G);
}

template <int NCH, int SPLIT, int PIX>
static int launch_gram_regs(const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean, double* gram_out,
                            cudaStream_t st) {
  const long long npix = (long long)(y1 - y0) * (x1 - x0);
  long long ctas = (npix + PIX - 1) / PIX;
  const long long cap = wctb_num_sms();          // one CTA per SM (register-limited), persistent over its pixels
  if (ctas > cap) ctas = cap;
  // iters = ceil(npix / stride), stride = ctas*PIX  =>  (iters-1)*stride < npix, so slot pbase + it*stride (pbase < stride)
  // is in range for every thread while it <= iters-2; only the last iteration is range-checked in the kernel.
  const long long stride = ctas * PIX;
  const unsigned iters = (unsigned)((npix + stride - 1) / stride);
  const unsigned wreg = (unsigned)(x1 - x0);
  if (x0 == 0 && x1 == W)
    gram_regs_kernel<NCH, SPLIT, PIX, true><<<(unsigned)ctas, PIX * SPLIT, 0, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                    (unsigned)npix, iters, mean, gram_out);
  else
    gram_regs_kernel<NCH, SPLIT, PIX, false><<<(unsigned)ctas, PIX * SPLIT, 0, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                     (unsigned)npix, iters, mean, gram_out);
  WCTB_RETURN_LAUNCH();
}

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

// ------------------------------------------------------------------------------------------
// Variant 4 (written after the round's last GPU slot: not yet run on hardware).  Same ring-fed register accumulation with
// TWO pixels per thread and iteration (pixels j and j + PIX of a 2*PIX-pixel tile, processed one after the other so the
// register budget is unchanged) and the range-checked last tile peeled out of the loop: half the barriers, copy-issue
// and loop overhead per pixel, and a loop body without the inlined checked copy -- the three stall sources ncu shows for
// variant 0 (profiles/r01_gram_ring_ncu_full.txt).  Kept as a separate copy so that the validated kernels stay
// byte-identical.
// ------------------------------------------------------------------------------------------
template <int NCH, int SPLIT, int PIX, bool FULLROW>
__device__ __forceinline__ void gram_ring2_issue(float4* __restrict__ ring_slot, const float4* __restrict__ x, long long HW,
                                                 int W, int y0, int x0, unsigned wreg, unsigned npix, unsigned tile_base) {
  constexpr int NT = PIX * SPLIT;
  constexpr int TP = 2 * PIX;                       // pixels per tile
#pragma unroll
  for (int e0 = 0; e0 < NCH * TP; e0 += NT) {
    const int e = e0 + (int)threadIdx.x;
    if (e < NCH * TP) {
      const int c = e / TP, jj = e - c * TP;
      const unsigned p = tile_base + jj;
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

template <int NCH, int SPLIT, int PART, int PIX, int RING, bool FULLROW>
__device__ __forceinline__ void gram_ring2_body(float4* __restrict__ ring, const float4* __restrict__ x, long long HW, int W, int y0,
                                                int x0, unsigned wreg, unsigned npix, unsigned iters,
                                                const float4* __restrict__ s_mh, double* __restrict__ G) {
  using D = GramDeal<NCH, SPLIT>;
  constexpr int C = NCH * 4;
  constexpr int NP = D::MAXCOUNT;
  constexpr int TP = 2 * PIX;
  constexpr int TILE = NCH * TP;    // float4 per ring slot
  float acc[NP][16];
#pragma unroll
  for (int s = 0; s < NP; ++s)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[s][e] = 0.f;
  const unsigned stride = gridDim.x * TP;
  const unsigned cta_base = blockIdx.x * TP;
  const int j = threadIdx.x % PIX;
#pragma unroll
  for (int t = 0; t < RING - 1; ++t) {
    if ((unsigned)t < iters) gram_ring2_issue<NCH, SPLIT, PIX, FULLROW>(ring + t * TILE, x, HW, W, y0, x0, wreg, npix, cta_base + t * stride);
    cp_async_commit();
  }
  for (unsigned it = 0; it + 1 < iters; ++it) {      // full tiles: every pixel slot is in range (see launch_gram_ring2)
    cp_async_wait<RING - 2>();
    asm volatile("bar.sync 1, %0;" ::"r"(PIX * SPLIT) : "memory");
    const unsigned tn = it + (RING - 1);
    if (tn < iters) gram_ring2_issue<NCH, SPLIT, PIX, FULLROW>(ring + (tn % RING) * TILE, x, HW, W, y0, x0, wreg, npix, cta_base + tn * stride);
    cp_async_commit();
    const float4* tile = ring + (it % RING) * TILE;
    // the step reads chunk c of pixel jpix at tile[c * PIXROW + jpix]; gram_ring_step uses PIX as the row length, so it is
    // instantiated with the tile's row length TP and called once per pixel
    gram_ring_step<NCH, SPLIT, PART, TP, false>(acc, tile, j, true, s_mh);
    gram_ring_step<NCH, SPLIT, PART, TP, false>(acc, tile, j + PIX, true, s_mh);
  }
  {
    const unsigned it = iters - 1;
    cp_async_wait<0>();
    asm volatile("bar.sync 1, %0;" ::"r"(PIX * SPLIT) : "memory");
    const float4* tile = ring + (it % RING) * TILE;
    const unsigned tb = cta_base + it * stride;
    gram_ring_step<NCH, SPLIT, PART, TP, true>(acc, tile, j, tb + j < npix, s_mh);
    gram_ring_step<NCH, SPLIT, PART, TP, true>(acc, tile, j + PIX, tb + j + PIX < npix, s_mh);
  }
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

template <int NCH, int SPLIT, int PART, int PIX, int RING, bool FULLROW>
__device__ __forceinline__ void gram_ring2_dispatch(int part, float4* __restrict__ ring, const float4* __restrict__ x, long long HW,
                                                    int W, int y0, int x0, unsigned wreg, unsigned npix, unsigned iters,
                                                    const float4* s_mh, double* __restrict__ G) {
  if (part == PART) {
    gram_ring2_body<NCH, SPLIT, PART, PIX, RING, FULLROW>(ring, x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  } else if constexpr (PART + 1 < SPLIT) {
    gram_ring2_dispatch<NCH, SPLIT, PART + 1, PIX, RING, FULLROW>(part, ring, x, HW, W, y0, x0, wreg, npix, iters, s_mh, G);
  }
}

template <int NCH, int SPLIT, int PIX, int RING, bool FULLROW>
__global__ void __launch_bounds__(PIX* SPLIT, 1) gram_ring2_kernel(const float4* __restrict__ x, int H, int W, int y0, int x0,
                                                                    unsigned wreg, unsigned npix, unsigned iters,
                                                                    const double* __restrict__ mean, double* __restrict__ G) {
  static_assert(PIX % 32 == 0, "a warp must not straddle two groups");
  static_assert(RING >= 3, "ring too shallow");
  extern __shared__ __align__(16) unsigned char ring_raw[];
  float4* ring = reinterpret_cast<float4*>(ring_raw);   // [RING][NCH][2*PIX]
  __shared__ float4 s_mh[NCH];
  if (threadIdx.x < NCH * 4) reinterpret_cast<float*>(s_mh)[threadIdx.x] = (float)mean[threadIdx.x];
  __syncthreads();
  gram_ring2_dispatch<NCH, SPLIT, 0, PIX, RING, FULLROW>(threadIdx.x / PIX, ring, x, (long long)H * W, W, y0, x0, wreg, npix, iters,
                                                         s_mh, G);
}

template <int NCH, int SPLIT, int PIX, int RING>
static int launch_gram_ring2(const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean, double* gram_out,
                             cudaStream_t st) {
  const long long npix = (long long)(y1 - y0) * (x1 - x0);
  constexpr int TP = 2 * PIX;
  long long ctas = (npix + TP - 1) / TP;
  const long long cap = wctb_num_sms();
  if (ctas > cap) ctas = cap;
  const long long stride = ctas * TP;
  const unsigned iters = (unsigned)((npix + stride - 1) / stride);   // (iters-1)*stride < npix: only the last tile is partial
  const unsigned wreg = (unsigned)(x1 - x0);
  const size_t smem = (size_t)RING * NCH * TP * sizeof(float4);
  static bool attr_done[WCTB_MAX_DEVICES] = {};      // per device: one process may drive several GPUs
  const int dev_slot = wctb_device_slot();
  if (!attr_done[dev_slot]) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(gram_ring2_kernel<NCH, SPLIT, PIX, RING, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WCTB_CUDA_TRY(cudaFuncSetAttribute(gram_ring2_kernel<NCH, SPLIT, PIX, RING, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev_slot] = true;
  }
  if (x0 == 0 && x1 == W)
    gram_ring2_kernel<NCH, SPLIT, PIX, RING, true><<<(unsigned)ctas, PIX * SPLIT, smem, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                              (unsigned)npix, iters, mean, gram_out);
  else
    gram_ring2_kernel<NCH, SPLIT, PIX, RING, false><<<(unsigned)ctas, PIX * SPLIT, smem, st>>>((const float4*)x, H, W, y0, x0, wreg,
                                                                                               (unsigned)npix, iters, mean, gram_out);
  WCTB_RETURN_LAUNCH();
}


int wctb_gram_regs_launch(int C, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                          double* gram_out, cudaStream_t st) {
  if (C == 24) return launch_gram_regs<6, 4, 96>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C == 32) return launch_gram_regs<8, 6, 64>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  return WCTB_E_UNSUPPORTED;
}

int wctb_gram_ring2_launch(int C, const float* x, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                           double* gram_out, cudaStream_t st) {
  if (C == 24) return launch_gram_ring2<6, 4, 96, 4>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C == 32) return launch_gram_ring2<8, 6, 64, 4>(x, H, W, y0, y1, x0, x1, mean, gram_out, st);
  return WCTB_E_UNSUPPORTED;
}
