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
#include "umma.cuh"

namespace {
using namespace wctb_umma;

// optional phase tracing (debug): clock64 stamps of a few CTAs, enabled through wctb_debug_set_trace
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int slot) {
  long long* t = g_trace;
  if (t && blockIdx.y == 10 && blockIdx.x < 8) t[blockIdx.x * 16 + slot] = clock64();
}

// ---------------------------------------------------------------------------------- geometry
template <int N> struct Cfg {
  // N <= 64: tiles sized so that TWO CTAs fit per SM (<= 256 TMEM columns, <= ~110 KB smem each): the epilogue of one
  // overlaps the MMAs / loads of the other.  N >= 128 (low-resolution layers): one CTA per SM, 512 columns.
  static constexpr int NB = N <= 32 ? 8 : (N == 64 ? 4 : (512 / N));   // accumulator blocks (128 positions each)
  static constexpr int TH = 2 * NB;                            // output rows per tile
  static constexpr int P = (TH + 2) * PW + 8;                  // pixel slots per chunk plane (+ slack for tap offsets)
  static constexpr int IN_BYTES = 2 * P * 16;                  // two 4-channel chunks
  static constexpr int W_BYTES = 9 * 2 * N * 16;               // [tap][chunk][N][4]
  static constexpr int STAGE_BYTES = IN_BYTES + W_BYTES;
  static constexpr int TMEM_COLS = (NB * N) < 32 ? 32 : (NB * N);
  static constexpr int POOL_BYTES = 2 * 64 * 20 * 4;           // double-buffered row-exchange for the pool epilogue
  static constexpr int AUX_BYTES = 1024;                       // barriers + tmem slot
  static constexpr int CTAS_PER_SM = (N <= 64) ? 2 : 1;
  static constexpr int SMEM_BUDGET = (CTAS_PER_SM == 2 ? 112 : 227) * 1024;
  static constexpr int NSTAGE_MAX = (SMEM_BUDGET - POOL_BYTES - AUX_BYTES - 128) / STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_MAX > 4 ? 4 : NSTAGE_MAX;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + POOL_BYTES + AUX_BYTES + 128;
  static_assert(NSTAGE >= 2, "pipeline needs two stages");
  static_assert(TMEM_COLS * CTAS_PER_SM <= 512, "TMEM budget");
};

struct ConvArgs {
  const float4* x;    // [Cin/4][H][W]
  const float* w;     // packed [nblk][Cin/8][9][2][N][4]
  const float* bias;  // [Cout]
  float4* y;
  int H, W, Cin, Cout, round_tf32;
};

// single-thread MMA issue loop: nkg pipeline stages x NB accumulator blocks x 9 taps (one K=8 MMA each)
template <int N, int NSTAGE, int STAGE_BYTES, int IN_BYTES, int P, int NB>
__device__ __forceinline__ void mma_issue_loop(uint8_t* stages, uint64_t* full, uint64_t* empty, uint64_t* accum_full,
                                               uint32_t tmem_base, int nkg) {
  constexpr uint32_t idesc = umma_idesc_tf32(N);
  for (int kg = 0; kg < nkg; ++kg) {
    const int slot = kg % NSTAGE;
    const uint32_t ph = (kg / NSTAGE) & 1;
    mbar_wait(full + slot, ph);
    tc_fence_after();
    const uint32_t a_base = smem_u32(stages + slot * STAGE_BYTES);
    const uint32_t w_base = a_base + IN_BYTES;
    // taps outer, accumulator blocks inner: consecutive MMAs hit DIFFERENT accumulators, so the ~80-cycle
    // accumulate-dependency latency of tcgen05.mma is hidden (measured: 80 -> ~10 cycles per N=16 MMA)
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap - dy * 3;
      const uint64_t bd = umma_desc(w_base + (uint32_t)tap * 2u * N * 16u, N * 16u, 128u);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint64_t ad = umma_desc(a_base + (uint32_t)(128 * b + dy * PW + dx) * 16u, P * 16u, 128u);
        umma_tf32(tmem_base + (uint32_t)(b * N), ad, bd, idesc, (kg > 0 || tap > 0) ? 1u : 0u);
      }
    }
    tc_commit(empty + slot);      // frees the smem stage when these MMAs have read it
  }
  tc_commit(accum_full);          // accumulators complete
}

// epilogue for the 4 warps owning the TMEM lane quarters: TMEM -> bias/ReLU/TF32-round -> (pool | up) -> P4 stores.
// Channel groups of 16 outer (bias in registers), accumulator blocks inner with the TMEM load of block b+1 in flight
// while block b is processed.
template <int N, int NB, int EPI>
__device__ __forceinline__ void conv_epilogue(const ConvArgs& a, uint64_t* accum_full, uint32_t tmem_base, float* poolbuf,
                                              int warp, int lane, int x0, int y0, int nblk) {
  const int H = a.H, W = a.W;
  const long long HW = (long long)H * W;
  const int q = warp & 3;           // TMEM lane quarter this warp may access
  mbar_wait(accum_full, 0);
  tc_fence_after();
  const float* bias = a.bias + nblk * N;
  const int cplane0 = nblk * (N / 4);
  const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16);
  const bool rnd = a.round_tf32 != 0;
  int it = 0;
#pragma unroll 1
  for (int g = 0; g < N / 16; ++g) {
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(bias + 16 * g + i);
    uint32_t buf[2][16];
    tmem_ld16_issue(tq + (uint32_t)(16 * g), buf[0]);
#pragma unroll
    for (int b = 0; b < NB; ++b, ++it) {
      uint32_t(&cur)[16] = buf[b & 1];
      tmem_ld16_wait(cur);
      if (b + 1 < NB) tmem_ld16_issue(tq + (uint32_t)((b + 1) * N + 16 * g), buf[(b + 1) & 1]);
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float t = wctb_relu(__uint_as_float(cur[i]) + bv[i]);
        v[i] = rnd ? wctb_tf32(t) : t;
      }
      if (EPI == WCTB_EPI_POOL2) {
        // block b = tile rows 2b (lanes 0..63) and 2b+1 (lanes 64..127): row exchange through shared memory
        const int Ho = H >> 1, Wo = W >> 1;
        const long long HWo = (long long)Ho * Wo;
        const int cpos = (q & 1) * 32 + lane;                    // column within the 64-pitch row
        const int oy = (y0 >> 1) + b, ox = (x0 + cpos) >> 1;
        const bool ok = (cpos < TW) && oy < Ho && ox < Wo && ((lane & 1) == 0);
        float* pb = poolbuf + (it & 1) * (64 * 20);
        if (q >= 2) {
          float4* d = reinterpret_cast<float4*>(pb + cpos * 20);
          d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
          d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (q < 2) {
          const float4* s = reinterpret_cast<const float4*>(pb + cpos * 20);
          const float4 s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
          const float o[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float m = fmaxf(v[i], o[i]);
            v[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          }
          if (ok) {
            const long long off = (long long)oy * Wo + ox;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              a.y[(long long)(cplane0 + 4 * g + j) * HWo + off] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
      } else {
        const int p = 128 * b + 32 * q + lane;
        const int r = p >> 6, c = p & 63;
        const int gy = y0 + r, gx = x0 + c;
        if ((c < TW) && gy < H && gx < W) {
          if (EPI == WCTB_EPI_NONE) {
            const long long off = (long long)gy * W + gx;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              a.y[(long long)(cplane0 + 4 * g + j) * HW + off] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else if (EPI == WCTB_EPI_NCHW3) {   // channels 0..2 of the (zero-padded) last layer as NCHW planes
            if (g == 0 && nblk == 0) {
              float* img = reinterpret_cast<float*>(a.y);
              const long long off = (long long)gy * W + gx;
              img[off] = v[0]; img[HW + off] = v[1]; img[2 * HW + off] = v[2];
            }
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

template <int N, int EPI>
__global__ void __launch_bounds__(192, Cfg<N>::CTAS_PER_SM) conv_umma_kernel(const ConvArgs a) {
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
    if (elect_one()) {
      for (int kg = 0; kg < nkg; ++kg) {
        const int slot = kg % C::NSTAGE;
        const uint32_t ph = (kg / C::NSTAGE) & 1;
        mbar_wait(empty + slot, ph ^ 1);
        uint8_t* st = stages + slot * C::STAGE_BYTES;
        mbar_expect_tx(full + slot, stage_tx);
        bulk_g2s(smem_u32(st + C::IN_BYTES), wblk + (size_t)kg * (C::W_BYTES / 4), C::W_BYTES, full + slot);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const float4* plane = a.x + (long long)(kg * 2 + c) * HW;
#pragma unroll 1
          for (int i = 0; i < C::TH + 2; ++i) {
            const int gy = wctb_reflect(y0 - 1 + i, H);
            const float4* src = plane + (long long)gy * W;
            const uint32_t dst = smem_u32(st + ((size_t)c * C::P + (size_t)i * PW) * 16);
            bulk_g2s(dst + jlo * 16, src + (x0 - 1 + jlo), (uint32_t)ncols * 16u, full + slot);
            if (left) bulk_g2s(dst, src + 1, 16u, full + slot);                          // gx = -1 -> 1
            if (right) bulk_g2s(dst + jr * 16, src + (W - 2), 16u, full + slot);         // gx = W  -> W-2
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one())
      mma_issue_loop<N, C::NSTAGE, C::STAGE_BYTES, C::IN_BYTES, C::P, C::NB>(stages, full, empty, accum_full, tmem_base, nkg);
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    conv_epilogue<N, C::NB, EPI>(a, accum_full, tmem_base, poolbuf, warp, lane, x0, y0, nblk);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int N, int EPI>
int launch_conv(const ConvArgs& a, cudaStream_t st) {
  using C = Cfg<N>;
  static bool attr_done[WCTB_MAX_DEVICES] = {};     // per device: one process may drive several GPUs
  WCTB_SET_SMEM_ONCE(attr_done, (conv_umma_kernel<N, EPI>), C::SMEM_BYTES);
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
    case WCTB_EPI_NCHW3: return N == 16 ? launch_conv<16, WCTB_EPI_NCHW3>(a, st) : WCTB_E_UNSUPPORTED;
    default: return launch_conv<N, WCTB_EPI_UP2>(a, st);
  }
}
inline int pick_n(int Cout) {
  if (Cout == 16 || Cout == 32 || Cout == 64 || Cout == 128 || Cout == 256) return Cout;
  if (Cout > 256 && Cout % 256 == 0) return 256;
  return 0;
}

// ---------------------------------------------------------------------------------- fused encoder head
//   y = [pool2]( ReLU(conv12( reflect_pad( ReLU(conv11( reflect_pad(img) )) ) )) )     img NCHW 3ch -> y P4 N ch
// conv11 (K = 27, not an MMA shape) is computed with FFMA by four producer warps straight into the tensor-core
// operand tile in shared memory (same [chunk][row][64] float4 layout the bulk-copy producer fills), 8 channels
// (= one pipeline stage) at a time; conv12 runs on tcgen05 from that tile.  The C1-channel full-resolution
// activation (64 B/px for the 16x net) never goes to HBM.  conv0 is folded into conv11 by the host.
// Reflection of the INTERMEDIATE: a halo position outside the image must hold conv11 evaluated at the mirrored
// position, so the producer evaluates conv11 at (reflect(y), reflect(x)).
struct HeadArgs {
  const float* img;   // [3][H][W]
  const float* w11;   // [27][C1]  (conv0 folded), fp32
  const float* b11;   // [C1]
  ConvArgs c;         // conv12: x unused, w packed tf32, bias, y, H, W, Cin = C1, Cout = N
};
template <int C1, int N> struct HeadCfg {
  using C = Cfg<N>;
  static constexpr int IMG_ROWS = C::TH + 4, IMG_PITCH = 68;
  static constexpr int IMG_BYTES = 3 * IMG_ROWS * IMG_PITCH * 4;
  static constexpr int W11_BYTES = (27 * C1 + C1) * 4;
  static constexpr int NSTAGE = 2;
  static constexpr int SMEM_BYTES = NSTAGE * C::STAGE_BYTES + IMG_BYTES + W11_BYTES + C::POOL_BYTES + 256 + 128;
  static constexpr int MIN_CTAS = (SMEM_BYTES <= 113 * 1024 && C::TMEM_COLS <= 256) ? 2 : 1;
};

template <int C1, int N, int EPI>
__global__ void __launch_bounds__(320, HeadCfg<C1, N>::MIN_CTAS) conv_head_kernel(const HeadArgs h) {
  using C = Cfg<N>;
  using HC = HeadCfg<C1, N>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* stages = smem;
  float* img_s = reinterpret_cast<float*>(smem + HC::NSTAGE * C::STAGE_BYTES);
  float* w11_s = reinterpret_cast<float*>(smem + HC::NSTAGE * C::STAGE_BYTES + HC::IMG_BYTES);
  float* b11_s = w11_s + 27 * C1;
  float* poolbuf = reinterpret_cast<float*>(smem + HC::NSTAGE * C::STAGE_BYTES + HC::IMG_BYTES + HC::W11_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(poolbuf) + C::POOL_BYTES);
  uint64_t* empty = full + HC::NSTAGE;
  uint64_t* accum_full = empty + HC::NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * C::TH;
  const int H = h.c.H, W = h.c.W;
  constexpr int nkg = C1 / KG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HC::NSTAGE; ++s) { mbar_init(full + s, 129); mbar_init(empty + s, 1); }
    mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- conv12 weight loader
    if (elect_one()) {
      for (int kg = 0; kg < nkg; ++kg) {
        const int slot = kg % HC::NSTAGE;
        mbar_wait(empty + slot, ((kg / HC::NSTAGE) & 1) ^ 1);
        mbar_expect_tx(full + slot, C::W_BYTES);
        bulk_g2s(smem_u32(stages + slot * C::STAGE_BYTES + C::IN_BYTES), h.c.w + (size_t)kg * (C::W_BYTES / 4), C::W_BYTES,
                 full + slot);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one())
      mma_issue_loop<N, HC::NSTAGE, C::STAGE_BYTES, C::IN_BYTES, C::P, C::NB>(stages, full, empty, accum_full, tmem_base, nkg);
    __syncwarp();
  } else if (warp < 6) {
    conv_epilogue<N, C::NB, EPI>(h.c, accum_full, tmem_base, poolbuf, warp, lane, x0, y0, 0);
  } else {
    // ---- conv11 producers (128 threads)
    const int tp = threadIdx.x - 192;
    const long long HW = (long long)H * W;
    for (int i = tp; i < 27 * C1 + C1; i += 128) w11_s[i] = i < 27 * C1 ? h.w11[i] : h.b11[i - 27 * C1];
    for (int i = tp; i < 3 * HC::IMG_ROWS * HC::IMG_PITCH; i += 128) {
      const int q = i % HC::IMG_PITCH, r = (i / HC::IMG_PITCH) % HC::IMG_ROWS, c = i / (HC::IMG_PITCH * HC::IMG_ROWS);
      const int gy = min(max(y0 - 2 + r, 0), H - 1), gx = min(max(x0 - 2 + q, 0), W - 1);
      img_s[i] = __ldg(h.img + c * HW + (long long)gy * W + gx);
    }
    asm volatile("bar.sync 2, 128;" ::: "memory");
    constexpr int NPAIR = (C::TH + 2) * PW / 2;
    for (int kg = 0; kg < nkg; ++kg) {
      const int slot = kg % HC::NSTAGE;
      mbar_wait(empty + slot, ((kg / HC::NSTAGE) & 1) ^ 1);
      float4* st0 = reinterpret_cast<float4*>(stages + slot * C::STAGE_BYTES);
      float4* st1 = st0 + C::P;
      for (int pr = tp; pr < NPAIR; pr += 128) {
        const int t = 2 * pr;                       // tile positions t, t+1 (same row)
        const int i = t / PW, j = t % PW;
        const int gy = wctb_reflect(y0 - 1 + i, H);
        int ry[3], cx[4];
#pragma unroll
        for (int d = 0; d < 3; ++d) ry[d] = min(max(wctb_reflect(gy + d - 1, H) - (y0 - 2), 0), HC::IMG_ROWS - 1) * HC::IMG_PITCH;
        const int gx0 = wctb_reflect(x0 - 1 + j, W), gx1 = wctb_reflect(x0 + j, W);
        float acc0[8], acc1[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) { acc0[o] = b11_s[kg * 8 + o]; acc1[o] = acc0[o]; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int c0 = min(max(wctb_reflect(gx0 + dx - 1, W) - (x0 - 2), 0), HC::IMG_PITCH - 1);
            const int c1 = min(max(wctb_reflect(gx1 + dx - 1, W) - (x0 - 2), 0), HC::IMG_PITCH - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float v0 = img_s[c * HC::IMG_ROWS * HC::IMG_PITCH + ry[dy] + c0];
              const float v1 = img_s[c * HC::IMG_ROWS * HC::IMG_PITCH + ry[dy] + c1];
              const float4* wp = reinterpret_cast<const float4*>(w11_s + ((dy * 3 + dx) * 3 + c) * C1 + kg * 8);
              const float4 wa = wp[0], wb = wp[1];
              acc0[0] = fmaf(v0, wa.x, acc0[0]); acc0[1] = fmaf(v0, wa.y, acc0[1]); acc0[2] = fmaf(v0, wa.z, acc0[2]); acc0[3] = fmaf(v0, wa.w, acc0[3]);
              acc0[4] = fmaf(v0, wb.x, acc0[4]); acc0[5] = fmaf(v0, wb.y, acc0[5]); acc0[6] = fmaf(v0, wb.z, acc0[6]); acc0[7] = fmaf(v0, wb.w, acc0[7]);
              acc1[0] = fmaf(v1, wa.x, acc1[0]); acc1[1] = fmaf(v1, wa.y, acc1[1]); acc1[2] = fmaf(v1, wa.z, acc1[2]); acc1[3] = fmaf(v1, wa.w, acc1[3]);
              acc1[4] = fmaf(v1, wb.x, acc1[4]); acc1[5] = fmaf(v1, wb.y, acc1[5]); acc1[6] = fmaf(v1, wb.z, acc1[6]); acc1[7] = fmaf(v1, wb.w, acc1[7]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) { acc0[o] = wctb_tf32(wctb_relu(acc0[o])); acc1[o] = wctb_tf32(wctb_relu(acc1[o])); }
        st0[t] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        st1[t] = make_float4(acc0[4], acc0[5], acc0[6], acc0[7]);
        st0[t + 1] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        st1[t + 1] = make_float4(acc1[4], acc1[5], acc1[6], acc1[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> visible to tcgen05.mma
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + slot)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int C1, int N, int EPI>
int launch_head(const HeadArgs& h, cudaStream_t st) {
  using C = Cfg<N>;
  using HC = HeadCfg<C1, N>;
  static bool attr_done[WCTB_MAX_DEVICES] = {};
  WCTB_SET_SMEM_ONCE(attr_done, (conv_head_kernel<C1, N, EPI>), HC::SMEM_BYTES);
  dim3 grid((h.c.W + TW - 1) / TW, (h.c.H + C::TH - 1) / C::TH, 1);
  if (grid.y > 65535) return WCTB_E_UNSUPPORTED;
  conv_head_kernel<C1, N, EPI><<<grid, 320, HC::SMEM_BYTES, st>>>(h);
  WCTB_RETURN_LAUNCH();
}

// ---------------------------------------------------------------------------------- fused encoder head, all-tensor-core (16x nets)
// Same fusion as conv_head_kernel, but conv11 (3 -> 16) also runs on tcgen05:
//  * the image tile is staged as RGB0 float4 pixels (P4 with C = 4) with pitch 68;
//  * with LBO = 16 bytes the second K-chunk of an MMA row is simply the NEXT PIXEL, so one K = 8 MMA covers the taps
//    (dy,dx) and (dy,dx+1): 3x3 taps = 6 MMAs per 128 positions (weights of the padding channel / 4th tap are zero);
//  * the conv11 accumulators are converted (bias, ReLU, TF32 round) into the two conv12 operand stages in shared
//    memory, halo positions outside the image are patched with the mirrored values, and conv12 + pool run as in the
//    generic kernel, re-using the same TMEM columns.
struct HeadTcArgs {
  const float* img;    // [3][H][W]
  const float* w11tc;  // [3 dy][2][2 chunks][16][4] tf32 (conv0 folded)
  const float* b11;    // [16]
  ConvArgs c;          // conv12
};
constexpr int HT_PI = 68;                 // image tile pitch (pixels)
constexpr int HT_IMG_ROWS = 21;           // 20 rows + 1 slack row for the tap offsets of garbage positions
constexpr int HT_NB1 = 10;                // conv11 accumulator blocks: 18 rows * 68 = 1224 positions
template <int EPI>
__global__ void __launch_bounds__(192, 2) conv_head_tc_kernel(const HeadTcArgs h) {
  constexpr int N = 16;
  using C = Cfg<N>;
  constexpr int IMG_BYTES = HT_IMG_ROWS * HT_PI * 16;      // 22848 (>= pool buffer, which aliases it later)
  constexpr int W11_BYTES = 6 * 2 * 16 * 16;               // 3072
  constexpr int TMEM_COLS = 256;
  static_assert(IMG_BYTES >= C::POOL_BYTES, "pool buffer aliases the image tile");
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* stages = smem;                                   // 2 x (operand tile + conv12 weights)
  float4* img_s = reinterpret_cast<float4*>(smem + 2 * C::STAGE_BYTES);
  float* poolbuf = reinterpret_cast<float*>(img_s);
  uint8_t* w11_s = smem + 2 * C::STAGE_BYTES + IMG_BYTES;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(w11_s + W11_BYTES);
  uint64_t* img_ready = wfull + 1;
  uint64_t* accum1_full = img_ready + 1;
  uint64_t* op_ready = accum1_full + 1;
  uint64_t* accum_full = op_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * C::TH;
  const int H = h.c.H, W = h.c.W;
  const long long HW = (long long)H * W;
  if (threadIdx.x == 0) trace(0);

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1); mbar_init(img_ready, 128); mbar_init(accum1_full, 1); mbar_init(op_ready, 128); mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace(1);

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(wfull, W11_BYTES + 2 * C::W_BYTES);
      bulk_g2s(smem_u32(w11_s), h.w11tc, W11_BYTES, wfull);
      bulk_g2s(smem_u32(stages + C::IN_BYTES), h.c.w, C::W_BYTES, wfull);
      bulk_g2s(smem_u32(stages + C::STAGE_BYTES + C::IN_BYTES), h.c.w + C::W_BYTES / 4, C::W_BYTES, wfull);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(N);
      mbar_wait(wfull, 0);
      mbar_wait(img_ready, 0);
      tc_fence_after();
      const uint32_t i_base = smem_u32(img_s), w1_base = smem_u32(w11_s);
#pragma unroll 1
      for (int t6 = 0; t6 < 6; ++t6) {
        const int dy = t6 >> 1, hh = t6 & 1;
        const uint64_t bd = umma_desc(w1_base + (uint32_t)t6 * 512u, 256u, 128u);
#pragma unroll
        for (int b = 0; b < HT_NB1; ++b) {
          const uint64_t ad = umma_desc(i_base + (uint32_t)(128 * b + dy * HT_PI + 2 * hh) * 16u, 16u, 128u);
          umma_tf32(tmem_base + (uint32_t)(b * N), ad, bd, idesc, t6 > 0 ? 1u : 0u);
        }
      }
      tc_commit(accum1_full);
      // ---- conv12 from the converted operand stages
      mbar_wait(op_ready, 0);
      tc_fence_after();
#pragma unroll 1
      for (int kg = 0; kg < 2; ++kg) {
        const uint32_t a_base = smem_u32(stages + kg * C::STAGE_BYTES);
        const uint32_t w_base = a_base + C::IN_BYTES;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint64_t bd = umma_desc(w_base + (uint32_t)tap * 2u * N * 16u, N * 16u, 128u);
#pragma unroll
          for (int b = 0; b < C::NB; ++b) {
            const uint64_t ad = umma_desc(a_base + (uint32_t)(128 * b + dy * PW + dx) * 16u, C::P * 16u, 128u);
            umma_tf32(tmem_base + (uint32_t)(b * N), ad, bd, idesc, (kg > 0 || tap > 0) ? 1u : 0u);
          }
        }
      }
      tc_commit(accum_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int et = threadIdx.x - 64;            // 0..127
    // ---- P0: RGB0 image tile (reflect-padded by 2; out-of-image conv11 positions are patched later)
    {
      constexpr int NIT = (HT_IMG_ROWS * HT_PI + 127) / 128;      // 12: all loads are issued before the first use
      float r0[NIT], r1[NIT], r2[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int idx = et + 128 * k;
        const int r = idx / HT_PI, qx = idx - r * HT_PI;
        r0[k] = r1[k] = r2[k] = 0.f;
        if (r < 20) {
          const int gy = wctb_reflect(y0 - 2 + r, H), gx = wctb_reflect(x0 - 2 + qx, W);
          const float* p = h.img + (long long)gy * W + gx;
          r0[k] = __ldg(p); r1[k] = __ldg(p + HW); r2[k] = __ldg(p + 2 * HW);
        }
      }
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int idx = et + 128 * k;
        if (idx < HT_IMG_ROWS * HT_PI) img_s[idx] = make_float4(wctb_tf32(r0[k]), wctb_tf32(r1[k]), wctb_tf32(r2[k]), 0.f);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(img_ready)) : "memory");
    if (et == 0) trace(2);
    // ---- E1: conv11 accumulators -> conv12 operand stages (pitch 64)
    float bb[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bb[i] = __ldg(h.b11 + i);
    float4* st0 = reinterpret_cast<float4*>(stages);
    float4* st1 = reinterpret_cast<float4*>(stages + C::STAGE_BYTES);
    mbar_wait(accum1_full, 0);
    tc_fence_after();
    if (et == 0) trace(3);
    for (int b = 0; b < HT_NB1; ++b) {
      const int p = 128 * b + 32 * q + lane;
      const int i = p / HT_PI, j = p - i * HT_PI;
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * N), v);
      if (i < C::TH + 2 && j < PW) {
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = wctb_tf32(wctb_relu(v[k] + bb[k]));
        const int d = i * PW + j;
        st0[d] = make_float4(v[0], v[1], v[2], v[3]);
        st0[C::P + d] = make_float4(v[4], v[5], v[6], v[7]);
        st1[d] = make_float4(v[8], v[9], v[10], v[11]);
        st1[C::P + d] = make_float4(v[12], v[13], v[14], v[15]);
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // reflection of the intermediate at true image borders (rows, then columns)
    const int rt = (y0 == 0) ? 0 : -1;
    const int rb = (H - y0 + 1 < C::TH + 2) ? (H - y0 + 1) : -1;
    if (rt == 0 || rb >= 2) {
      for (int e = et; e < 4 * PW; e += 128) {
        const int c = e & 63, pl = e >> 6;
        float4* base = (pl < 2 ? st0 : st1) + (pl & 1) * C::P;
        if (rt == 0) base[c] = base[2 * PW + c];
        if (rb >= 2) base[rb * PW + c] = base[(rb - 2) * PW + c];
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int cl = (x0 == 0) ? 0 : -1;
    const int cr = (W - x0 + 1 < PW) ? (W - x0 + 1) : -1;
    if ((cl == 0 || cr >= 2) && et < 4 * (C::TH + 2)) {
      const int r = et % (C::TH + 2), pl = et / (C::TH + 2);
      float4* base = (pl < 2 ? st0 : st1) + (pl & 1) * C::P;
      if (cl == 0) base[r * PW] = base[r * PW + 2];
      if (cr >= 2) base[r * PW + cr] = base[r * PW + cr - 2];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(op_ready)) : "memory");
    if (et == 0) trace(4);
    mbar_wait(accum_full, 0);
    if (et == 0) trace(5);
    // ---- E2: conv12 epilogue (pool buffer aliases the image tile, which is dead by now)
    conv_epilogue<N, C::NB, EPI>(h.c, accum_full, tmem_base, poolbuf, warp, lane, x0, y0, 0);
    if (et == 0) trace(6);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(7);
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int EPI>
int launch_head_tc(const HeadTcArgs& h, cudaStream_t st) {
  using C = Cfg<16>;
  constexpr int SMEM = 2 * C::STAGE_BYTES + HT_IMG_ROWS * HT_PI * 16 + 6 * 2 * 16 * 16 + 256 + 128;
  static bool attr_done[WCTB_MAX_DEVICES] = {};
  WCTB_SET_SMEM_ONCE(attr_done, (conv_head_tc_kernel<EPI>), SMEM);
  dim3 grid((h.c.W + TW - 1) / TW, (h.c.H + C::TH - 1) / C::TH, 1);
  if (grid.y > 65535) return WCTB_E_UNSUPPORTED;
  conv_head_tc_kernel<EPI><<<grid, 192, SMEM, st>>>(h);
  WCTB_RETURN_LAUNCH();
}

// ---------------------------------------------------------------------------------- fused decoder tail
//   img = ReLU(conv11( reflect_pad( ReLU(conv12( reflect_pad(x) )) ) ))      x P4 16ch (full res) -> img NCHW 3ch
// Both convolutions run on tcgen05.  conv12 (16->16) is computed exactly like the generic kernel, for an
// intermediate tile one pixel larger on every side than the 14x60 final tile.  Its epilogue keeps the ReLU'd,
// TF32-rounded intermediate in shared memory, in the same [chunk][row*64+col] float4 layout -- i.e. it is already a
// valid A operand -- patches the halo positions outside the image with the mirrored intermediate values (reflection
// applies to the INTERMEDIATE, not to conv12's input), and conv11 (16->3, zero-padded to N = 16) is a second round of
// MMAs over that tile into a second TMEM range; a final epilogue writes NCHW.  The 16-channel full-resolution
// conv12 output (64 B/px write + read) never goes to HBM.
// UPSRC: the input is the half-resolution tensor and nearest-x2 upsampling (model_cd.py:261) is applied while
// producer warps fill the operand tile (generic loads), so the upsampled tensor never exists in HBM either.
struct TailArgs {
  ConvArgs c;         // conv12: x (P4 16ch; [H/2][W/2] when UPSRC), w packed tf32, bias; y unused; H, W = OUTPUT size
  const float* w11;   // conv11 zero-padded to 16 outputs, packed tf32 like any 16->16 layer
  const float* b11;   // [3]
  float* img;         // [3][H][W]
};
constexpr int TAIL_TH = 14, TAIL_TW = 60;
constexpr int TAIL_IBP = 16 * PW + 8;      // intermediate plane pitch (float4) incl. slack for the tap offsets
template <bool UPSRC>
__global__ void __launch_bounds__(UPSRC ? 320 : 192, 2) conv_tail_kernel(const TailArgs t) {
  constexpr int N = 16;
  using C = Cfg<N>;            // NB = 8 -> 16 intermediate rows x 62 valid intermediate columns
  constexpr int NSTAGE = 2;
  constexpr int NB2 = TAIL_TH * PW / 128;    // 7 accumulator blocks for conv11
  constexpr int TMEM_COLS = 256;             // [0,128) conv12, [128,240) conv11
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* stages = smem;
  float4* ibuf = reinterpret_cast<float4*>(smem);                       // [4 chunks][TAIL_IBP] -- reuses the stages
  uint8_t* w11_s = smem + NSTAGE * C::STAGE_BYTES;                      // 2 * W_BYTES: [kg][tap][2][16][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(w11_s + 2 * C::W_BYTES);
  uint64_t* empty = full + NSTAGE;
  uint64_t* accum_full = empty + NSTAGE;
  uint64_t* w11_full = accum_full + 1;
  uint64_t* ibuf_ready = w11_full + 1;
  uint64_t* accum2_full = ibuf_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum2_full + 1);
  static_assert(4 * TAIL_IBP * 16 <= NSTAGE * C::STAGE_BYTES, "intermediate tile must fit in the freed stages");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x0f = blockIdx.x * TAIL_TW, y0f = blockIdx.y * TAIL_TH;     // final tile origin
  const int x0 = x0f - 1, y0 = y0f - 1;                                 // intermediate tile origin
  const int H = t.c.H, W = t.c.W;
  const long long HW = (long long)H * W;
  constexpr int nkg = 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, UPSRC ? 129 : 1); mbar_init(empty + s, 1); }
    mbar_init(accum_full, 1);
    mbar_init(w11_full, 1);
    mbar_init(ibuf_ready, 128);
    mbar_init(accum2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer: weights always by bulk copy; input rows by bulk copy unless UPSRC
    const bool elected = elect_one();
    if (elected) {
      mbar_expect_tx(w11_full, 2 * C::W_BYTES);
      bulk_g2s(smem_u32(w11_s), t.w11, 2 * C::W_BYTES, w11_full);
    }
    if (!UPSRC) {
      // intermediate tile col j <-> input col gx = x0 - 1 + j (x0 may be -1)
      const int jlo = max(0, 1 - x0);                                     // first tile col with gx >= 0
      const int jhi = min(PW, W - x0 + 1);
      const int ncols = max(jhi - jlo, 0);
      const uint32_t nrefl_l = (uint32_t)jlo;                             // gx = -2, -1 -> 2, 1
      const int jr = W - x0 + 1;                                          // tile col of gx == W
      const uint32_t nrefl_r = (uint32_t)min(max(PW - jr, 0), 2);         // gx = W, W+1 -> W-2, W-3
      const uint32_t row_bytes = ((uint32_t)ncols + nrefl_l + nrefl_r) * 16u;
      const uint32_t stage_tx = (uint32_t)C::W_BYTES + 2u * (C::TH + 2) * row_bytes;
      if (elected) {
        for (int kg = 0; kg < nkg; ++kg) {
          uint8_t* st = stages + kg * C::STAGE_BYTES;
          mbar_expect_tx(full + kg, stage_tx);
          bulk_g2s(smem_u32(st + C::IN_BYTES), t.c.w + (size_t)kg * (C::W_BYTES / 4), C::W_BYTES, full + kg);
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const float4* plane = t.c.x + (long long)(kg * 2 + c) * HW;
#pragma unroll 1
            for (int i = 0; i < C::TH + 2; ++i) {
              const int gy = wctb_reflect(y0 - 1 + i, H);
              const float4* src = plane + (long long)gy * W;
              const uint32_t dst = smem_u32(st + ((size_t)c * C::P + (size_t)i * PW) * 16);
              if (ncols > 0) bulk_g2s(dst + jlo * 16, src + (x0 - 1 + jlo), (uint32_t)ncols * 16u, full + kg);
              for (uint32_t k = 0; k < nrefl_l; ++k) bulk_g2s(dst + k * 16, src + wctb_reflect(x0 - 1 + (int)k, W), 16u, full + kg);
              for (uint32_t k = 0; k < nrefl_r; ++k) bulk_g2s(dst + (jr + k) * 16, src + wctb_reflect(W + (int)k, W), 16u, full + kg);
            }
          }
        }
      }
    } else if (elected) {
      for (int kg = 0; kg < nkg; ++kg) {
        mbar_expect_tx(full + kg, C::W_BYTES);
        bulk_g2s(smem_u32(stages + kg * C::STAGE_BYTES + C::IN_BYTES), t.c.w + (size_t)kg * (C::W_BYTES / 4), C::W_BYTES, full + kg);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      mma_issue_loop<N, NSTAGE, C::STAGE_BYTES, C::IN_BYTES, C::P, C::NB>(stages, full, empty, accum_full, tmem_base, nkg);
      // ---- conv11 over the intermediate tile
      mbar_wait(w11_full, 0);
      mbar_wait(ibuf_ready, 0);
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_tf32(N);
      const uint32_t a_base = smem_u32(ibuf), w_base = smem_u32(w11_s);
#pragma unroll 1
      for (int kt = 0; kt < 18; ++kt) {
        const int kg = kt / 9, tap = kt - kg * 9;
        const int dy = tap / 3, dx = tap - dy * 3;
        const uint64_t bd = umma_desc(w_base + (uint32_t)(kt * 2) * N * 16u, N * 16u, 128u);
#pragma unroll
        for (int b = 0; b < NB2; ++b) {
          const uint64_t ad = umma_desc(a_base + (uint32_t)(2 * kg * TAIL_IBP + 128 * b + dy * PW + dx) * 16u, TAIL_IBP * 16u, 128u);
          umma_tf32(tmem_base + 128u + (uint32_t)(b * N), ad, bd, idesc, kt > 0 ? 1u : 0u);
        }
      }
      tc_commit(accum2_full);
    }
    __syncwarp();
  } else if (warp < 6) {
    // ---- epilogue 1: intermediate -> smem (TF32), border fix-up;  epilogue 2: conv11 accumulators -> NCHW
    const int q = warp & 3;
    const int et = threadIdx.x - 64;            // 0..127
    mbar_wait(accum_full, 0);
    tc_fence_after();
    for (int b = 0; b < C::NB; ++b) {
      const int p = 128 * b + 32 * q + lane;
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b * N), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = wctb_tf32(wctb_relu(v[i] + __ldg(t.c.bias + i)));
#pragma unroll
      for (int j = 0; j < 4; ++j) ibuf[j * TAIL_IBP + p] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // reflection of the intermediate at true image borders: rows first, then columns (corners follow)
    const int rt = (y0 < 0) ? 0 : -1;                         // tile row holding global row -1
    const int rb = (H - y0 < 16) ? (H - y0) : -1;             // tile row holding global row H
    if (rt == 0 || rb >= 2) {
      for (int i = et; i < 4 * PW; i += 128) {
        const int c = i & 63, ch = i >> 6;
        if (rt == 0) ibuf[ch * TAIL_IBP + c] = ibuf[ch * TAIL_IBP + 2 * PW + c];
        if (rb >= 2) ibuf[ch * TAIL_IBP + rb * PW + c] = ibuf[ch * TAIL_IBP + (rb - 2) * PW + c];
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int cl = (x0 < 0) ? 0 : -1;
    const int cr = (W - x0 < PW) ? (W - x0) : -1;
    if ((cl == 0 || cr >= 2) && et < 4 * 16) {
      const int r = et & 15, ch = et >> 4;
      if (cl == 0) ibuf[ch * TAIL_IBP + r * PW] = ibuf[ch * TAIL_IBP + r * PW + 2];
      if (cr >= 2) ibuf[ch * TAIL_IBP + r * PW + cr] = ibuf[ch * TAIL_IBP + r * PW + cr - 2];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> visible to tcgen05.mma
    tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(ibuf_ready)) : "memory");
    // ---- final epilogue
    const float b0 = __ldg(t.b11), b1 = __ldg(t.b11 + 1), b2 = __ldg(t.b11 + 2);
    mbar_wait(accum2_full, 0);
    tc_fence_after();
    for (int b = 0; b < NB2; ++b) {
      const int p = 128 * b + 32 * q + lane;
      const int r = p >> 6, c = p & 63;
      float v[4];
      tmem_ld4(tmem_base + ((uint32_t)(32 * q) << 16) + 128u + (uint32_t)(b * N), v);
      const int gy = y0f + r, gx = x0f + c;
      if (c < TAIL_TW && gy < H && gx < W) {
        const long long o = (long long)gy * W + gx;
        t.img[o] = wctb_relu(v[0] + b0);
        t.img[HW + o] = wctb_relu(v[1] + b1);
        t.img[2 * HW + o] = wctb_relu(v[2] + b2);
      }
    }
  } else if (UPSRC) {
    // ---- upsampling producers (128 threads): half-res source -> full-res operand tile
    const int tp = threadIdx.x - 192;
    const int Hs = H >> 1, Ws = W >> 1;
    const long long HWs = (long long)Hs * Ws;
    for (int kg = 0; kg < nkg; ++kg) {
      float4* st = reinterpret_cast<float4*>(stages + kg * C::STAGE_BYTES);
      constexpr int NEL = 2 * (C::TH + 2) * PW;        // 2304 = 18 per thread, in 3 batches of 6 loads in flight
#pragma unroll 1
      for (int base = 0; base < NEL; base += 6 * 128) {
        float4 v[6];
        int dsto[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int idx = base + tp + 128 * k;
          const int j = idx & 63, i = (idx >> 6) % (C::TH + 2), c = idx / (PW * (C::TH + 2));
          const int gy = wctb_reflect(y0 - 1 + i, H), gx = wctb_reflect(x0 - 1 + j, W);
          dsto[k] = c * C::P + i * PW + j;
          v[k] = __ldg(t.c.x + (long long)(kg * 2 + c) * HWs + (long long)(gy >> 1) * Ws + (gx >> 1));
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) st[dsto[k]] = v[k];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + kg)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <bool UPSRC>
int launch_tail(const TailArgs& t, cudaStream_t st) {
  using C = Cfg<16>;
  constexpr int SMEM = 2 * C::STAGE_BYTES + 2 * C::W_BYTES + 256 + 128;
  static bool attr_done[WCTB_MAX_DEVICES] = {};
  WCTB_SET_SMEM_ONCE(attr_done, (conv_tail_kernel<UPSRC>), SMEM);
  dim3 grid((t.c.W + TAIL_TW - 1) / TAIL_TW, (t.c.H + TAIL_TH - 1) / TAIL_TH, 1);
  if (grid.y > 65535) return WCTB_E_UNSUPPORTED;
  conv_tail_kernel<UPSRC><<<grid, UPSRC ? 320 : 192, SMEM, st>>>(t);
  WCTB_RETURN_LAUNCH();
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

// ---------------------------------------------------------------------------------- MMA issue-rate microbenchmark (debug)
// cycles per tcgen05.mma (M=128, K=8 tf32) for a given N and operand layout: 0 = SWIZZLE_NONE planes (what the conv
// kernels use), 1 = SWIZZLE_128B rows (128 B per row, K advanced by 32 B inside the row), 2 = SWIZZLE_NONE with
// LBO = 16 (the tap-pair trick).  Operand contents are irrelevant (uninitialised smem).
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int layout, int nacc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
    constexpr uint32_t idesc = umma_idesc_tf32(N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (layout == 1) {   // SWIZZLE_128B K-major: SBO = 1024, layout_type = 2
          ad = (uint64_t)(((a0 + k * 32) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          bd = (uint64_t)(((b0 + k * 32) >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        } else if (layout == 2) {
          ad = umma_desc(a0 + k * 1088, 16u, 128u);
          bd = umma_desc(b0 + k * 2 * N * 16, N * 16u, 128u);
        } else {
          ad = umma_desc(a0 + k * 64 * 16, 18560u, 128u);
          bd = umma_desc(b0 + k * 2 * N * 16, N * 16u, 128u);
        }
#pragma unroll 1
        for (int b = 0; b < nacc; ++b) umma_tf32(tmem + (uint32_t)(b * N), ad, bd, idesc, 1u);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
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

extern "C" int wctb_conv_head_supported(int C1, int Cout) { return ((C1 == 16 && Cout == 16) || (C1 == 64 && Cout == 64)) ? 1 : 0; }

extern "C" int wctb_conv_head(const float* x_nchw, const float* w11, const float* b11, const float* w12_packed,
                              const float* b12, float* y_p4, int H, int W, int C1, int Cout, int epilogue, int round_tf32,
                              void* stream) {
  if (!x_nchw || !w11 || !b11 || !w12_packed || !b12 || !y_p4 || H < 2 || W < 2) return WCTB_E_BADARG;
  if (epilogue != WCTB_EPI_NONE && epilogue != WCTB_EPI_POOL2) return WCTB_E_BADARG;
  if (!wctb_conv_head_supported(C1, Cout)) return WCTB_E_UNSUPPORTED;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  HeadArgs h{x_nchw, w11, b11, ConvArgs{nullptr, w12_packed, b12, (float4*)y_p4, H, W, C1, Cout, round_tf32}};
  cudaStream_t st = (cudaStream_t)stream;
  if (C1 == 16) return epilogue == WCTB_EPI_POOL2 ? launch_head<16, 16, WCTB_EPI_POOL2>(h, st) : launch_head<16, 16, WCTB_EPI_NONE>(h, st);
  return epilogue == WCTB_EPI_POOL2 ? launch_head<64, 64, WCTB_EPI_POOL2>(h, st) : launch_head<64, 64, WCTB_EPI_NONE>(h, st);
}

extern "C" int wctb_conv_tail_supported(int Cin, int Cmid) { return (Cin == 16 && Cmid == 16) ? 1 : 0; }

extern "C" int wctb_conv_tail(const float* x_p4, const float* w12_packed, const float* b12, const float* w11,
                              const float* b11, float* y_nchw, int H, int W, int Cin, int Cmid, int upsample_input,
                              void* stream) {
  if (!x_p4 || !w12_packed || !b12 || !w11 || !b11 || !y_nchw || H < 2 || W < 2) return WCTB_E_BADARG;
  if (!wctb_conv_tail_supported(Cin, Cmid)) return WCTB_E_UNSUPPORTED;
  if (upsample_input && ((H & 1) || (W & 1))) return WCTB_E_BADARG;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  TailArgs t{ConvArgs{(const float4*)x_p4, w12_packed, b12, nullptr, H, W, Cin, Cmid, 0}, w11, b11, y_nchw};
  return upsample_input ? launch_tail<true>(t, (cudaStream_t)stream) : launch_tail<false>(t, (cudaStream_t)stream);
}

extern "C" int wctb_conv_head_tc(const float* x_nchw, const float* w11_tc, const float* b11, const float* w12_packed,
                                 const float* b12, float* y_p4, int H, int W, int epilogue, int round_tf32, void* stream) {
  if (!x_nchw || !w11_tc || !b11 || !w12_packed || !b12 || !y_p4 || H < 2 || W < 2) return WCTB_E_BADARG;
  if (epilogue != WCTB_EPI_NONE && epilogue != WCTB_EPI_POOL2) return WCTB_E_BADARG;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  HeadTcArgs h{x_nchw, w11_tc, b11, ConvArgs{nullptr, w12_packed, b12, (float4*)y_p4, H, W, 16, 16, round_tf32}};
  return epilogue == WCTB_EPI_POOL2 ? launch_head_tc<WCTB_EPI_POOL2>(h, (cudaStream_t)stream)
                                    : launch_head_tc<WCTB_EPI_NONE>(h, (cudaStream_t)stream);
}

extern "C" int wctb_debug_set_trace(long long* buf) {
  WCTB_CUDA_TRY(cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
  return WCTB_OK;
}

extern "C" int wctb_debug_mma_rate(long long* out_cycles, int N, int layout, int nacc, int iters, int ctas, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 200 * 1024;
#define WCTB_MR(NN) case NN: WCTB_CUDA_TRY(cudaFuncSetAttribute(mma_rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    mma_rate_kernel<NN><<<ctas, 128, smem, st>>>(out_cycles, layout, nacc, iters); break;
  switch (N) { WCTB_MR(16) WCTB_MR(32) WCTB_MR(64) WCTB_MR(128) WCTB_MR(256) default: return WCTB_E_UNSUPPORTED; }
#undef WCTB_MR
  WCTB_RETURN_LAUNCH();
}
