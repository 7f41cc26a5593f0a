// libwctb: tcgen05 (5th-gen tensor core) TF32 implicit-GEMM convolution engine for sm_100a.
//
//   y = [pool2 | up2]( ReLU( conv3x3( reflect_pad1(x) ) + bias ) )      x, y in P4 layout [C/4][H][W][4] fp32
//
// Mapping.  GEMM M = output pixels, N = output channels, K = 9 taps x Cin.  A CTA owns an output tile of
// TH = 2*NB rows x 62 columns.  The input halo tile lives in shared memory with a row pitch of PW = 64
// pixels, one plane per 4-channel chunk:   tile[chunk][row][col] of float4   (== the P4 layout itself).
// That is exactly the tcgen05 "K-major, no-swizzle" canonical operand layout: a core matrix is 8 rows
// (= 8 consecutive pixels, 16 B apart) x 16 B (= 4 channels); the next 8-row group is +128 B (SBO), the
// next 4 channels are +plane stride (LBO).  Because the tile is linear in memory, the operand of filter tap
// (dy,dx) for the 128 consecutive tile positions [128 b, 128 b + 128) is the SAME buffer at byte offset
// ((128 b + dy*64 + dx) * 16): nine taps = nine descriptor start addresses, no im2col copies.  The two
// rightmost positions of each 64-pixel row are garbage outputs and are masked in the epilogue (3%).
//   - weights are pre-packed per 8-channel K-group and tap as [chunk][Cout][4] (K-major B operand),
//     pre-rounded to TF32 (rna); activations are rounded by their producer's epilogue, so the tensor core's
//     operand truncation is exact.
//   - accumulators: NB blocks of 128 x N fp32 in TMEM (NB*N <= 512 columns).
// Pipeline per CTA (192 threads): warp 0 = producer (cp.async.bulk row copies, reflection resolved in the
// source address; completion on an mbarrier), warp 1 = TMEM allocator + single-thread MMA issuer
// (tcgen05.mma.cta_group::1.kind::tf32, tcgen05.commit frees the stage), warps 2..5 = epilogue
// (tcgen05.ld -> bias/ReLU/TF32-round -> pool|up -> coalesced float4 stores).
#include "common.cuh"

namespace {

constexpr int PW = 64;   // smem row pitch (pixels)
constexpr int TW = 62;   // valid output columns per tile
constexpr int KG = 8;    // channels per pipeline stage (one K=8 MMA per tap)

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done, addr = smem_u32(bar);
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(NCOLS) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 (K-direction core-matrix stride) | [32,46) SBO>>4 (M/N-direction 8-row stride)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------- geometry
template <int N> struct Cfg {
  static constexpr int NB = (512 / N) < 8 ? (512 / N) : 8;     // accumulator blocks (128 positions each)
  static constexpr int TH = 2 * NB;                            // output rows per tile
  static constexpr int P = (TH + 2) * PW + 8;                  // pixel slots per chunk plane (+ slack for tap offsets)
  static constexpr int IN_BYTES = 2 * P * 16;                  // two 4-channel chunks
  static constexpr int W_BYTES = 9 * 2 * N * 16;               // [tap][chunk][N][4]
  static constexpr int STAGE_BYTES = IN_BYTES + W_BYTES;
  static constexpr int TMEM_COLS = (NB * N) < 32 ? 32 : (NB * N);
  static constexpr int POOL_BYTES = 2 * 64 * 20 * 4;           // double-buffered row-exchange for the pool epilogue
  static constexpr int AUX_BYTES = 1024;                       // barriers + tmem slot
  static constexpr int NSTAGE_MAX = (227 * 1024 - POOL_BYTES - AUX_BYTES - 128) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_MAX > 4 ? 4 : NSTAGE_MAX;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + POOL_BYTES + AUX_BYTES + 128;
  static_assert(NSTAGE >= 2, "pipeline needs two stages");
};

struct ConvArgs {
  const float4* x;    // [Cin/4][H][W]
  const float* w;     // packed [nblk][Cin/8][9][2][N][4]
  const float* bias;  // [Cout]
  float4* y;
  int H, W, Cin, Cout, round_tf32;
};

template <int N, int EPI>
__global__ void __launch_bounds__(192, 1) conv_umma_kernel(const ConvArgs a) {
  using C = Cfg<N>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* stages = smem;
  float* poolbuf = reinterpret_cast<float*>(smem + C::NSTAGE * C::STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE_BYTES + C::POOL_BYTES);
  uint64_t* empty = full + C::NSTAGE;
  uint64_t* accum_full = empty + C::NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * C::TH;
  const int nblk = blockIdx.z;                    // output-channel block of N
  const int H = a.H, W = a.W;
  const long long HW = (long long)H * W;
  const int nkg = a.Cin / KG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== producer ===========================
    const int jlo = (x0 == 0) ? 1 : 0;                                   // tile col j <-> gx = x0 - 1 + j
    const int jhi = min(PW, W - x0 + 1);                                 // exclusive
    const int ncols = jhi - jlo;
    const bool left = (x0 == 0);
    const int jr = W - x0 + 1;                                           // tile col of gx == W (right reflection)
    const bool right = jr < PW;
    const uint32_t row_bytes = (uint32_t)(ncols + (left ? 1 : 0) + (right ? 1 : 0)) * 16u;
    const uint32_t stage_tx = (uint32_t)C::W_BYTES + 2u * (C::TH + 2) * row_bytes;
    const float* wblk = a.w + (size_t)nblk * nkg * (C::W_BYTES / 4);
    for (int kg = 0; kg < nkg; ++kg) {
      const int slot = kg % C::NSTAGE;
      const uint32_t ph = (kg / C::NSTAGE) & 1;
      mbar_wait(empty + slot, ph ^ 1);
      uint8_t* st = stages + slot * C::STAGE_BYTES;
      if (lane == 0) {
        mbar_expect_tx(full + slot, stage_tx);
        bulk_g2s(smem_u32(st + C::IN_BYTES), wblk + (size_t)kg * (C::W_BYTES / 4), C::W_BYTES, full + slot);
      }
      __syncwarp();
      for (int idx = lane; idx < 2 * (C::TH + 2); idx += 32) {
        const int c = idx / (C::TH + 2), i = idx - c * (C::TH + 2);
        const int gy = wctb_reflect(y0 - 1 + i, H);
        const float4* src = a.x + (long long)(kg * 2 + c) * HW + (long long)gy * W;
        const uint32_t dst = smem_u32(st + ((size_t)c * C::P + (size_t)i * PW) * 16);
        bulk_g2s(dst + jlo * 16, src + (x0 - 1 + jlo), (uint32_t)ncols * 16u, full + slot);
        if (left) bulk_g2s(dst, src + 1, 16u, full + slot);                          // gx = -1 -> 1
        if (right) bulk_g2s(dst + jr * 16, src + (W - 2), 16u, full + slot);         // gx = W  -> W-2
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(N);
      for (int kg = 0; kg < nkg; ++kg) {
        const int slot = kg % C::NSTAGE;
        const uint32_t ph = (kg / C::NSTAGE) & 1;
        mbar_wait(full + slot, ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(stages + slot * C::STAGE_BYTES);
        const uint32_t w_base = a_base + C::IN_BYTES;
#pragma unroll 1
        for (int b = 0; b < C::NB; ++b) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint64_t ad = umma_desc(a_base + (uint32_t)(128 * b + dy * PW + dx) * 16u, C::P * 16u, 128u);
            const uint64_t bd = umma_desc(w_base + (uint32_t)tap * 2u * N * 16u, N * 16u, 128u);
            umma_tf32(tmem_base + (uint32_t)(b * N), ad, bd, idesc, (kg > 0 || tap > 0) ? 1u : 0u);
          }
        }
        tc_commit(empty + slot);      // frees the smem stage when these MMAs have read it
      }
      tc_commit(accum_full);          // accumulators complete
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const float* bias = a.bias + nblk * N;
    const int cplane0 = nblk * (N / 4);
    if (EPI == WCTB_EPI_POOL2) {
      const int Ho = H >> 1, Wo = W >> 1;
      const long long HWo = (long long)Ho * Wo;
      int it = 0;
      for (int b = 0; b < C::NB; ++b) {
        // block b = tile rows 2b (lanes 0..63) and 2b+1 (lanes 64..127)
        const int cpos = (q & 1) * 32 + lane;                    // column within the 64-pitch row
        const int oy = (y0 >> 1) + b, ox = (x0 + cpos) >> 1;
        const bool ok = (cpos < TW) && oy < Ho && ox < Wo && ((lane & 1) == 0);
        for (int g = 0; g < N / 16; ++g, ++it) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * N + 16 * g), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t = wctb_relu(v[i] + __ldg(bias + 16 * g + i));
            v[i] = a.round_tf32 ? wctb_tf32(t) : t;
          }
          float* buf = poolbuf + (it & 1) * (64 * 20);
          if (q >= 2) {
            float4* d = reinterpret_cast<float4*>(buf + cpos * 20);
            d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
            d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (q < 2) {
            const float4* s = reinterpret_cast<const float4*>(buf + cpos * 20);
            float4 s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
            float o[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float m = fmaxf(v[i], o[i]);
              v[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            }
            if (ok) {
              const long long off = (long long)oy * Wo + ox;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a.y[(long long)(cplane0 + 4 * g + j) * HWo + off] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          }
        }
      }
    } else {
      for (int b = 0; b < C::NB; ++b) {
        const int p = 128 * b + 32 * q + lane;
        const int r = p >> 6, c = p & 63;
        const int gy = y0 + r, gx = x0 + c;
        const bool ok = (c < TW) && gy < H && gx < W;
        for (int g = 0; g < N / 16; ++g) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * N + 16 * g), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float t = wctb_relu(v[i] + __ldg(bias + 16 * g + i));
            v[i] = a.round_tf32 ? wctb_tf32(t) : t;
          }
          if (ok) {
            if (EPI == WCTB_EPI_NONE) {
              const long long off = (long long)gy * W + gx;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a.y[(long long)(cplane0 + 4 * g + j) * HW + off] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {  // nearest x2
              const int Wo = 2 * W;
              const long long HWo = 4 * HW;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                float4* pl = a.y + (long long)(cplane0 + 4 * g + j) * HWo;
                const long long off = (long long)(2 * gy) * Wo + 2 * gx;
                pl[off] = o; pl[off + 1] = o; pl[off + Wo] = o; pl[off + Wo + 1] = o;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int N, int EPI>
int launch_conv(const ConvArgs& a, cudaStream_t st) {
  using C = Cfg<N>;
  static bool attr_done = false;
  if (!attr_done) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(conv_umma_kernel<N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dim3 grid((a.W + TW - 1) / TW, (a.H + C::TH - 1) / C::TH, a.Cout / N);
  if (grid.y > 65535 || grid.z > 65535) return WCTB_E_UNSUPPORTED;
  conv_umma_kernel<N, EPI><<<grid, 192, C::SMEM_BYTES, st>>>(a);
  WCTB_RETURN_LAUNCH();
}
template <int N>
int launch_conv_epi(const ConvArgs& a, int epi, cudaStream_t st) {
  switch (epi) {
    case WCTB_EPI_NONE: return launch_conv<N, WCTB_EPI_NONE>(a, st);
    case WCTB_EPI_POOL2: return launch_conv<N, WCTB_EPI_POOL2>(a, st);
    default: return launch_conv<N, WCTB_EPI_UP2>(a, st);
  }
}
inline int pick_n(int Cout) {
  if (Cout == 16 || Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256) return Cout;
  if (Cout > 256 && Cout % 256 == 0) return 256;
  return 0;
}

// ---------------------------------------------------------------------------------- weight packing
__global__ void pack_w_tf32_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cin, int Cout, int N) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = 9LL * Cin * Cout;
  if (i >= total) return;
  // dst index = ((((nb*nkg + kg)*9 + tap)*2 + c)*N + n)*4 + e
  int e = (int)(i & 3);
  long long t = i >> 2;
  int n = (int)(t % N); t /= N;
  int c = (int)(t & 1); t >>= 1;
  int tap = (int)(t % 9); t /= 9;
  int nkg = Cin / KG;
  int kg = (int)(t % nkg);
  int nb = (int)(t / nkg);
  int co = nb * N + n, ci = kg * KG + c * 4 + e;
  dst[i] = wctb_tf32(w[((long long)co * Cin + ci) * 9 + tap]);
}

// ---------------------------------------------------------------------------------- self test: D[128xN] = A[128xK] B[NxK]^T
template <int N>
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(float* __restrict__ out, const float* __restrict__ A,
                                                               const float* __restrict__ B, int K) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* sa = reinterpret_cast<float*>(smem);                 // [K/4][128][4]
  float* sb = sa + (size_t)K * 128;                           // [K/4][N][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + (size_t)K * N);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * K; i += 128) {
    int m = i / K, k = i - m * K;
    sa[((k >> 2) * 128 + m) * 4 + (k & 3)] = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += 128) {
    int n = i / K, k = i - n * K;
    sb[((k >> 2) * N + n) * 4 + (k & 3)] = B[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr int COLS = N < 32 ? 32 : N;
  if (warp == 0) tmem_alloc<COLS>(slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    for (int k8 = 0; k8 < K / 8; ++k8) {
      uint64_t ad = umma_desc(smem_u32(sa) + (uint32_t)k8 * 2u * 128u * 16u, 128u * 16u, 128u);
      uint64_t bd = umma_desc(smem_u32(sb) + (uint32_t)k8 * 2u * N * 16u, N * 16u, 128u);
      umma_tf32(tmem, ad, bd, umma_idesc_tf32(N), k8 > 0 ? 1u : 0u);
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int g = 0; g < N / 16; ++g) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + 16 * g, v);
    for (int i = 0; i < 16; ++i) out[(size_t)(32 * warp + lane) * N + 16 * g + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<COLS>(tmem);
}

}  // namespace

int wctb_conv3x3_p4_tf32_impl(const float* x, const float* w, const float* bias, float* y, int H, int W, int Cin,
                              int Cout, int epilogue, int round_tf32, cudaStream_t st) {
  const int N = pick_n(Cout);
  if (N == 0 || (Cin % KG) != 0) return WCTB_E_UNSUPPORTED;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  ConvArgs a{(const float4*)x, w, bias, (float4*)y, H, W, Cin, Cout, round_tf32};
  switch (N) {
    case 16: return launch_conv_epi<16>(a, epilogue, st);
    case 32: return launch_conv_epi<32>(a, epilogue, st);
    case 64: return launch_conv_epi<64>(a, epilogue, st);
    case 128: return launch_conv_epi<128>(a, epilogue, st);
    default: return launch_conv_epi<256>(a, epilogue, st);
  }
}

extern "C" int wctb_tf32_supported(int Cin, int Cout) { return (pick_n(Cout) != 0 && Cin > 0 && (Cin % KG) == 0) ? 1 : 0; }
extern "C" int wctb_tf32_kgroup(int Cin, int Cout) { return wctb_tf32_supported(Cin, Cout) ? KG : 0; }

extern "C" int wctb_pack_weights_tf32(const float* w, float* dst, int Cin, int Cout, void* stream) {
  if (!w || !dst) return WCTB_E_BADARG;
  if (!wctb_tf32_supported(Cin, Cout)) return WCTB_E_UNSUPPORTED;
  long long total = 9LL * Cin * Cout;
  pack_w_tf32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, dst, Cin, Cout, pick_n(Cout));
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_selftest_umma(float* out, const float* a, const float* b, int N, int K, void* stream) {
  if (!out || !a || !b || K <= 0 || (K % 8) != 0 || K > 64) return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)K * (128 + N) * 4 + 256;
#define WCTB_ST(NN)                                                                                                   \
  case NN:                                                                                                            \
    WCTB_CUDA_TRY(cudaFuncSetAttribute(umma_selftest_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    umma_selftest_kernel<NN><<<1, 128, smem, st>>>(out, a, b, K);                                                     \
    break;
  switch (N) {
    WCTB_ST(16) WCTB_ST(32) WCTB_ST(64) WCTB_ST(128) WCTB_ST(256)
    default: return WCTB_E_UNSUPPORTED;
  }
#undef WCTB_ST
  WCTB_RETURN_LAUNCH();
}
