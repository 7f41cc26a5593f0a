// libwctb: "h2" convolution engine -- fp32-accurate 3x3 convolutions on the 5th-gen tensor cores (sm_100a).
//
//   y = [pool2 | up2]( ReLU( conv3x3( reflect_pad1(x) ) + bias ) )
//
// Why.  Single-pass TF32 (10-bit mantissa operands) does not meet the precision contract of the path: chained through
// the five whiten/colour stages its operand rounding is amplified to 2.1e-2 RMS on the images bench.py feeds
// (tools/precision_probe.py).  Here every fp32 operand is carried as a PAIR of fp16 numbers, x = hi + lo with
// hi = fp16(x), lo = fp16(x - hi) (22 significand bits), and a product is evaluated as hi*w_hi + lo*w_hi + hi*w_lo with
// tcgen05.mma.kind::f16 into fp32 TMEM accumulators (the dropped lo*w_lo term is 2^-22 relative).  Weights are scaled
// per layer by a power of two (device side, no host sync) so that w_lo stays in the fp16 normal range; the epilogue
// multiplies by the inverse.  kind::f16 has K = 16 per instruction against K = 8 for kind::tf32 and the N <= 64 layers
// are bound by the A-operand fetch (one 4 KB slab per MMA whatever N is), so the pair costs the same number of MMAs
// as single-pass TF32 when the two weight halves are stacked along N:
//      D[:, 0:N)  += A_hi * W_hi^T + A_lo * W_hi^T          D[:, N:2N) += A_hi * W_lo^T        (2 MMAs per tap / 16 ch)
// and the HBM footprint of an activation is unchanged (2 + 2 bytes per element).
//
// Activation layout "H8": [C/8 chunks][2 (hi, lo)][H][W][8] fp16 -- one 16-byte unit = 8 consecutive channels of a
// pixel, i.e. already the tcgen05 K-major / no-swizzle core-matrix row, so a rectangular halo tile copied into shared
// memory IS the A operand and the nine filter taps are nine descriptor start addresses (linear-pitch trick, row pitch
// 64 pixels, the two rightmost positions of a row are masked garbage).
//
// Kernel structure (persistent, one CTA per SM, 320 threads):
//   warp 0   producer: per 16-channel K group ONE tensor-map TMA box (cp.async.bulk.tensor.3d, 4 planes x (TH+2) rows x
//            1 KB = 64 px) + one bulk copy of the weight slab (or the whole layer's weights once, H2Cfg::RESW); tiles touching a
//            true image border are copied by the 32 lanes with the reflection resolved in the source index.
//   warp 1   single-thread tcgen05.mma issuer; accumulators double-buffered in TMEM (2 x NB blocks of 128 px), so the
//            MMAs of tile t+1 overlap the epilogue of tile t inside the CTA.
//   warps 2-9 epilogue (two groups of four, alternating accumulator blocks): tcgen05.ld -> (main + minor) * 1/s + bias -> ReLU
//            -> [pool | up2] -> hi/lo split -> 16-byte stores (H8 for the next conv and/or fp32 P4 for the statistics kernels /
//            public API).
#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "h2.cuh"

namespace {
using namespace wctb_umma;

// ---------------------------------------------------------------------------------- geometry
template <int N_, int NB_, int STACK_, int RESW_ = 0>
struct H2Cfg {
  static constexpr int N = N_, NB = NB_, STACK = STACK_;
  // RESW: the packed weights of the whole layer (<= MAXKG 16-channel K groups, Cout == N) stay RESIDENT in shared memory for the
  // life of the persistent CTA instead of being re-fetched with every pipeline stage.  Why: for N <= 64 the tcgen05.mma rate is
  // bound by the 128 B/clk shared-memory read port (cycles per MMA = (4 KB A slab + 32 N bytes of B) / 128: 40 / 48 / 64 for
  // N = 32 / 64 / 128, tools/h2_rates.py), the TMA writes of a stage share that port with the operand reads, and the weight slab
  // was 30-47 % of every stage (ncu: L2 -> SM reads 3.4-3.8x the input tensor).
  static constexpr int RESW = RESW_;
  static constexpr int MAXKG = 4;
  static constexpr int NACC = 2;                                 // accumulator sets (double buffer)
  static constexpr int TH = 2 * NB;                              // output rows per tile
  static constexpr int ROWS = TH + 2;
  static constexpr int PLANE_BYTES = ROWS * PW * 16;             // one 8-channel half-plane of the halo tile
  static constexpr int IN_BYTES = 4 * PLANE_BYTES;               // hi0, lo0, hi1, lo1
  static constexpr int W_BYTES = 9 * 2 * (2 * N) * 16;           // [tap][kchunk][hi N rows | lo N rows][8 halves]
  static constexpr int STAGE_BYTES = IN_BYTES + (RESW ? 0 : W_BYTES);
  static constexpr int RES_BYTES = RESW ? MAXKG * W_BYTES : 0;
  static constexpr int COLS = STACK ? 2 * N : N;                 // TMEM columns per accumulator block
  static constexpr int ACC_COLS = NB * COLS;
  static constexpr int AUX_BYTES = 512;
  static_assert(NACC * ACC_COLS <= 512, "TMEM budget");
  static_assert(STAGE_BYTES % 128 == 0 && W_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(COLS <= 256, "tcgen05.mma N <= 256");
};
// shared-memory plan of one (config, epilogue) instantiation: [resident weights][stages][pool exchange][barriers]
template <class C, int EPI>
struct H2Lay {
  static constexpr int POOL_BYTES = (EPI == WCTB_EPI_POOL2) ? 2 * 2 * 64 * 20 * 4 : 0;   // [warp group][parity][64 px][16 + 4 pad]
  static constexpr int NSTAGE_MAX = (227 * 1024 - C::RES_BYTES - POOL_BYTES - C::AUX_BYTES - 128) / C::STAGE_BYTES;
  static constexpr int NSTAGE = NSTAGE_MAX > 4 ? 4 : NSTAGE_MAX;
  static constexpr int STAGES_OFF = C::RES_BYTES;
  static constexpr int POOL_OFF = STAGES_OFF + NSTAGE * C::STAGE_BYTES;
  static constexpr int AUX_OFF = POOL_OFF + POOL_BYTES;
  static constexpr int SMEM_BYTES = AUX_OFF + C::AUX_BYTES + 128;
  static_assert(NSTAGE >= 2, "pipeline needs two stages");
  static_assert((2 * NSTAGE + 2 * C::NACC + C::MAXKG) * 8 + 8 <= C::AUX_BYTES, "barrier area");
};

constexpr int H2_EPI_WARPS = 8;                     // two warp groups of four (one warp per TMEM lane quarter each)
constexpr int H2_THREADS = 64 + 32 * H2_EPI_WARPS;   // warp 0 producer, warp 1 MMA issuer, warps 2.. epilogue

struct H2Args {
  const __half* x;        // H8 [C8in][2][H][W][8]
  const __half* w;        // packed, see pack_w_h2_kernel
  const float* bias;      // [Cout]
  const float* wscale;    // device: [1] = 1 / (power-of-two weight scale)
  __half* y_h8;           // H8 output (nullable)
  float4* y_p4;           // fp32 P4 output (nullable); EPI_NCHW3: the [3][H][W] image
  int H, W, Cin, Cout;
  int tiles_x, tiles_y, ntiles;   // ntiles = nblks * tiles_y * tiles_x
  int planes_in;                  // 2 * ceil(Cin / 8)
};

// ---------------------------------------------------------------------------------- MMA issue (one elected thread)
// One pipeline stage = 16 input channels: planes [hi0, lo0, hi1, lo1] (K chunk stride = 2 planes) + weight slab.
// taps outer / accumulator blocks inner, and the MMAs that accumulate into the same columns are issued in separate
// passes over the blocks, so dependent tcgen05.mma are >= NB instructions apart (accumulate latency ~90 cycles).
template <class C>
__device__ __forceinline__ void h2_issue_stage(uint32_t a_base, uint32_t w_base, uint32_t tmem_acc, bool first,
                                               int a_pitch = PW, uint32_t a_lbo = 2u * C::PLANE_BYTES,
                                               uint32_t a_lo_off = C::PLANE_BYTES) {
  constexpr int N = C::N;
  constexpr uint32_t id_n = umma_idesc_f16(N), id_2n = umma_idesc_f16(2 * N);
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;
    const uint32_t acc = (first && tap == 0) ? 0u : 1u;
    const uint32_t wb = w_base + (uint32_t)tap * (2u * 2u * N * 16u);
    const uint64_t bd = umma_desc(wb, 2u * N * 16u, 128u);                    // rows [0,2N): hi | lo
    const uint32_t aoff = (uint32_t)(dy * a_pitch + dx) * 16u;
    if (C::STACK) {
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
        umma_f16(tmem_acc + (uint32_t)(b * C::COLS), umma_desc(a_base + aoff + (uint32_t)b * 2048u, a_lbo, 128u), bd, id_2n, acc);
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
        umma_f16(tmem_acc + (uint32_t)(b * C::COLS), umma_desc(a_base + a_lo_off + aoff + (uint32_t)b * 2048u, a_lbo, 128u), bd, id_n, 1u);
    } else {
      const uint64_t bd_lo = umma_desc(wb + (uint32_t)N * 16u, 2u * N * 16u, 128u);
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
        umma_f16(tmem_acc + (uint32_t)(b * C::COLS), umma_desc(a_base + aoff + (uint32_t)b * 2048u, a_lbo, 128u), bd, id_n, acc);
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
        umma_f16(tmem_acc + (uint32_t)(b * C::COLS), umma_desc(a_base + a_lo_off + aoff + (uint32_t)b * 2048u, a_lbo, 128u), bd, id_n, 1u);
#pragma unroll
      for (int b = 0; b < C::NB; ++b)
        umma_f16(tmem_acc + (uint32_t)(b * C::COLS), umma_desc(a_base + aoff + (uint32_t)b * 2048u, a_lbo, 128u), bd_lo, id_n, 1u);
    }
  }
}

// ---------------------------------------------------------------------------------- epilogue
// store 16 consecutive channels (global 16-channel group gg) of one pixel
__device__ __forceinline__ void h2_store16(const H2Args& a, int gg, long long HWo, long long off, const float* v) {
  if (a.y_h8) {
    uint4* base = reinterpret_cast<uint4*>(a.y_h8);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint4 hi, lo;
      split8(v + 8 * j, hi, lo);
      const long long pl = (long long)(2 * gg + j) * 2;
      base[pl * HWo + off] = hi;
      base[(pl + 1) * HWo + off] = lo;
    }
  }
  if (a.y_p4) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      a.y_p4[(long long)(4 * gg + j) * HWo + off] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}

// One accumulator item = 16 output channels (group g) of one 128-pixel block b: bias, ReLU, [pool | up2], hi/lo split, stores.
template <class C, int EPI>
__device__ __forceinline__ void h2_epilogue_item(const H2Args& a, const uint32_t (&m0)[16], const uint32_t (&m1)[16], const float (&bv)[16],
                                                 float* poolbuf, int half, int q, int lane, int x0, int y0, int nblk, int g, int b,
                                                 float inv_s, int& it) {
  constexpr int N = C::N;
  const int H = a.H, W = a.W;
  const long long HW = (long long)H * W;
  const int gg = nblk * (N / 16) + g;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float s = C::STACK ? (__uint_as_float(m0[i]) + __uint_as_float(m1[i])) : __uint_as_float(m0[i]);
    v[i] = wctb_relu(fmaf(s, inv_s, bv[i]));
  }
  if (EPI == WCTB_EPI_POOL2) {
    // block b = tile rows 2b (lanes 0..63) and 2b+1 (lanes 64..127): row exchange through shared memory (per warp group)
    const int Ho = H >> 1, Wo = W >> 1;
    const int cpos = (q & 1) * 32 + lane;
    const int oy = (y0 >> 1) + b, ox = (x0 + cpos) >> 1;
    const bool ok = (cpos < TW) && oy < Ho && ox < Wo && ((lane & 1) == 0);
    float* pb = poolbuf + (half * 2 + (it & 1)) * (64 * 20);
    if (q >= 2) {
      float4* d = reinterpret_cast<float4*>(pb + cpos * 20);
      d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
      d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
    if (q < 2) {
      const float4* s = reinterpret_cast<const float4*>(pb + cpos * 20);
      const float4 s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
      const float o[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float m = fmaxf(v[i], o[i]);
        v[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      }
      if (ok) h2_store16(a, gg, (long long)Ho * Wo, (long long)oy * Wo + ox, v);
    }
  } else {
    const int p = 128 * b + 32 * q + lane;
    const int r = p >> 6, c = p & 63;
    const int gy = y0 + r, gx = x0 + c;
    if ((c < TW) && gy < H && gx < W) {
      if (EPI == WCTB_EPI_NONE) {
        h2_store16(a, gg, HW, (long long)gy * W + gx, v);
      } else if (EPI == WCTB_EPI_NCHW3) {
        if (g == 0 && nblk == 0) {
          float* img = reinterpret_cast<float*>(a.y_p4);
          const long long off = (long long)gy * W + gx;
          img[off] = v[0]; img[HW + off] = v[1]; img[2 * HW + off] = v[2];
        }
      } else {   // nearest x2
        const int Wo = 2 * W;
        const long long off = (long long)(2 * gy) * Wo + 2 * gx;
        if (a.y_h8) {
          uint4* base = reinterpret_cast<uint4*>(a.y_h8);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4 hi, lo;
            split8(v + 8 * j, hi, lo);
            uint4* ph = base + (long long)(2 * gg + j) * 2 * (4 * HW);
            uint4* pl = ph + 4 * HW;
            ph[off] = hi; ph[off + 1] = hi; ph[off + Wo] = hi; ph[off + Wo + 1] = hi;
            pl[off] = lo; pl[off + 1] = lo; pl[off + Wo] = lo; pl[off + Wo + 1] = lo;
          }
        }
        if (a.y_p4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            float4* pl = a.y_p4 + (long long)(4 * gg + j) * (4 * HW);
            pl[off] = o; pl[off + 1] = o; pl[off + Wo] = o; pl[off + Wo + 1] = o;
          }
        }
      }
    }
  }
  ++it;
}

// The epilogue of one tile, as seen by ONE of the two warp groups (half = 0 / 1: four warps each, one per TMEM lane quarter).
// A group takes the blocks b = half (mod 2) of every 16-channel group: with a single warp per SM sub-partition the epilogue
// was bound by its own dependent-instruction latency (tcgen05.ld -> math -> stores, ~500 cycles per item at 13 % issue
// utilisation) and, not the MMAs, set the tile time of every layer up to 64 channels; two warps per sub-partition halve it.
// Items are walked g-major with the TMEM load of item i+1 in flight while item i is processed.
template <class C, int EPI>
__device__ __forceinline__ void h2_epilogue_tile(const H2Args& a, uint32_t tmem_acc, float* poolbuf, int half, int q, int lane,
                                                 int x0, int y0, int nblk, float inv_s, int& it) {
  constexpr int N = C::N, BH = C::NB / 2, ITEMS = (N / 16) * BH;
  static_assert(C::NB % 2 == 0 && ITEMS % 2 == 0, "two warp groups, items processed in pairs");
  const uint32_t tq = tmem_acc + ((uint32_t)(32 * q) << 16);
  auto issue = [&](int i, uint32_t (&r0)[16], uint32_t (&r1)[16]) {
    const int g = i / BH, b = 2 * (i - g * BH) + half;
    tmem_ld16_issue(tq + (uint32_t)(b * C::COLS + 16 * g), r0);
    if (C::STACK) tmem_ld16_issue(tq + (uint32_t)(b * C::COLS + N + 16 * g), r1);
  };
  float bv[16];
  int g_loaded = -1;
  auto run = [&](int i, uint32_t (&r0)[16], uint32_t (&r1)[16]) {
    const int g = i / BH, b = 2 * (i - g * BH) + half;
    if (g != g_loaded) {
#pragma unroll
      for (int k = 0; k < 16; ++k) bv[k] = __ldg(a.bias + nblk * N + 16 * g + k);
      g_loaded = g;
    }
    h2_epilogue_item<C, EPI>(a, r0, r1, bv, poolbuf, half, q, lane, x0, y0, nblk, g, b, inv_s, it);
  };
  uint32_t ma0[16], ma1[16], mb0[16], mb1[16];
  issue(0, ma0, ma1);
#pragma unroll 1
  for (int i = 0; i < ITEMS; i += 2) {
    tmem_ld16_wait(ma0);
    if (C::STACK) tmem_ld16_wait(ma1);
    issue(i + 1, mb0, mb1);
    run(i, ma0, ma1);
    tmem_ld16_wait(mb0);
    if (C::STACK) tmem_ld16_wait(mb1);
    if (i + 2 < ITEMS) issue(i + 2, ma0, ma1);
    run(i + 1, mb0, mb1);
  }
}

// ---------------------------------------------------------------------------------- producer helpers
// Border tile (one that touches a true image edge): the whole producer warp copies the (TH+2) rows x 4 planes x 64 px halo
// tile with 16-byte cp.async (all of them in flight at once), the reflection resolved per pixel in the source index (columns
// and rows beyond the image are clamped: they only feed masked outputs).  The stage is handed to the MMA issuer one stage
// later (cp.async.wait_group + proxy fence + arrive), so consecutive border stages overlap.  Round 2a issued one cp.async.bulk per row and per reflected
// 16-byte column from a single thread -- 50-70 small bulk copies per stage, ~7 us, six times an interior stage; with 14 %
// (960x540) to 56 % (240x135) of the tiles on a border that, not the MMAs, set the kernel time (tools/profile_h2_generic.py:
// interior tile 4.6 us, border tile 28.8 us).
template <class C>
__device__ __forceinline__ void h2_load_border_warp(const H2Args& a, uint8_t* st, int kg, int x0, int y0, int lane) {
  const int H = a.H, W = a.W;
  const long long HW = (long long)H * W;
  const uint4* xb = reinterpret_cast<const uint4*>(a.x);
  const int gxa = wctb_reflect(x0 - 1 + lane, W), gxb = wctb_reflect(x0 - 1 + 32 + lane, W);
  const uint32_t sbase = smem_u32(st) + (uint32_t)lane * 16u;
#pragma unroll 1
  for (int pl = 0; pl < 4; ++pl) {
    int gp = 4 * kg + pl;
    if (gp >= a.planes_in) gp -= 2;                  // Cin % 16 == 8: the missing chunk has zero weights; feed finite data
    const uint4* plane = xb + (long long)gp * HW;
    const uint32_t sp = sbase + (uint32_t)(pl * C::ROWS) * (PW * 16u);
#pragma unroll
    for (int i = 0; i < C::ROWS; ++i) {
      const uint4* src = plane + (long long)wctb_reflect(y0 - 1 + i, H) * W;
      cpa16(sp + (uint32_t)i * (PW * 16u), src + gxa);
      cpa16(sp + (uint32_t)i * (PW * 16u) + 512u, src + gxb);
    }
  }
  cpa_commit();                                      // one group per stage: the caller completes it one stage later
}

template <class C, int EPI>
__global__ void __launch_bounds__(H2_THREADS, 1) conv_h2_kernel(const __grid_constant__ CUtensorMap tmap, const H2Args a) {
  using L = H2Lay<C, EPI>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* resw = smem;                                         // RESW: [nkg][W_BYTES]
  uint8_t* stages = smem + L::STAGES_OFF;
  float* poolbuf = reinterpret_cast<float*>(smem + L::POOL_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::AUX_OFF);
  uint64_t* empty = full + L::NSTAGE;
  uint64_t* acc_full = empty + L::NSTAGE;
  uint64_t* acc_empty = acc_full + C::NACC;
  uint64_t* wfull = acc_empty + C::NACC;                        // RESW: one per K group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + C::MAXKG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkg = (a.Cin + 15) / 16;
  const int tiles_xy = a.tiles_x * a.tiles_y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < L::NSTAGE; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < C::NACC; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, H2_EPI_WARPS); }
    for (int s = 0; s < C::MAXKG; ++s) mbar_init(wfull + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmap);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== producer (whole warp; one elected lane drives the TMA unit) ===========================
    if (C::RESW) {                         // the layer's weights, once per CTA (K group by K group: the first MMAs need only slab 0)
      if (elect_one()) {
        for (int kg = 0; kg < nkg; ++kg) {
          mbar_expect_tx(wfull + kg, (uint32_t)C::W_BYTES);
          bulk_g2s(smem_u32(resw + (size_t)kg * C::W_BYTES), a.w + (size_t)kg * (C::W_BYTES / 2), C::W_BYTES, wfull + kg);
        }
      }
      __syncwarp();
    }
    uint32_t it = 0;
    int pend = -1;                         // slot of a border stage whose cp.async group has not been handed over yet
    auto hand_over = [&](auto keep) {      // complete all but the `keep` most recent cp.async groups and publish the pending stage
      if (pend >= 0) {
        cpa_wait<decltype(keep)::value>();
        fence_async_smem();                // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (elect_one()) mbar_arrive(full + pend);
        __syncwarp();
        pend = -1;
      }
    };
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const int nblk = tile / tiles_xy, rem = tile - nblk * tiles_xy;
      const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
      const int x0 = tx * TW, y0 = ty * C::TH;
      const bool border = (x0 == 0) || (y0 == 0) || (x0 + TW >= a.W) || (y0 + C::TH >= a.H);
      const __half* wblk = a.w + (size_t)nblk * nkg * (C::W_BYTES / 2);
      for (int kg = 0; kg < nkg; ++kg, ++it) {
        const int slot = it % L::NSTAGE;
        uint8_t* st = stages + slot * C::STAGE_BYTES;
        const __half* wsrc = wblk + (size_t)kg * (C::W_BYTES / 2);
        if (!border) {
          hand_over(std::integral_constant<int, 0>{});
          if (elect_one()) {
            mbar_wait(empty + slot, ((it / L::NSTAGE) & 1) ^ 1);
            mbar_expect_tx(full + slot, (uint32_t)C::STAGE_BYTES);
            tma_load_3d(smem_u32(st), &tmap, 4 * (x0 - 1), y0 - 1, 4 * kg, full + slot);   // planes beyond the tensor: zero fill
            if (!C::RESW) bulk_g2s(smem_u32(st + C::IN_BYTES), wsrc, C::W_BYTES, full + slot);
          }
          __syncwarp();
        } else {
          mbar_wait(empty + slot, ((it / L::NSTAGE) & 1) ^ 1);
          if (!C::RESW) {
            if (elect_one()) {
              mbar_expect_tx_only(full + slot, (uint32_t)C::W_BYTES);
              bulk_g2s(smem_u32(st + C::IN_BYTES), wsrc, C::W_BYTES, full + slot);
            }
            __syncwarp();
          }
          h2_load_border_warp<C>(a, st, kg, x0, y0, lane);
          hand_over(std::integral_constant<int, 1>{});      // the previous border stage (this one stays in flight)
          pend = slot;
        }
      }
    }
    hand_over(std::integral_constant<int, 0>{});
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t % C::NACC;
        mbar_wait(acc_empty + acc, ((t / C::NACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc * C::ACC_COLS;
        for (int kg = 0; kg < nkg; ++kg, ++it) {
          const int slot = it % L::NSTAGE;
          if (C::RESW && t == 0) mbar_wait(wfull + kg, 0);
          mbar_wait(full + slot, (it / L::NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(stages + slot * C::STAGE_BYTES);
          const uint32_t w_base = C::RESW ? smem_u32(resw + (size_t)kg * C::W_BYTES) : a_base + C::IN_BYTES;
          h2_issue_stage<C>(a_base, w_base, tmem_acc, kg == 0);
          tc_commit(empty + slot);
        }
        tc_commit(acc_full + acc);
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const float inv_s = __ldg(a.wscale + 1);
    uint32_t t = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++t) {
      const int nblk = tile / tiles_xy, rem = tile - nblk * tiles_xy;
      const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
      const uint32_t acc = t % C::NACC;
      mbar_wait(acc_full + acc, (t / C::NACC) & 1);
      tc_fence_after();
      h2_epilogue_tile<C, EPI>(a, tmem_base + acc * C::ACC_COLS, poolbuf, half, q, lane, tx * TW, ty * C::TH, nblk, inv_s, it);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// H8 tensor [planes][H][W][8] fp16 as a 3-D tensor map of 32-bit words: {W*4 words, H, planes}, box = 256 words (64 px x 16 B) x
// rows x 4 planes.  The innermost box extent is what the TMA unit moves per request: described as {8 halves, W, H, planes} (the
// natural 4-D view, inner extent 16 bytes) a 24 KB box became 1536 sixteen-byte requests and took ~4.2 us to land whatever
// its size -- that latency, not the MMAs or the epilogue, set the tile time of every layer up to 64 channels.  As rows of 1 KB
// it is 24 requests.
int make_h8_tmap(CUtensorMap* m, const void* base, int planes, int H, int W, int rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return WCTB_E_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)W * 4, (cuuint64_t)H, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {(cuuint32_t)PW * 4, (cuuint32_t)rows, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_wctb_last_cuda_error = (int)r; return WCTB_E_CUDA; }
  return WCTB_OK;
}

template <class C, int EPI>
int launch_h2(H2Args a, cudaStream_t st) {
  static bool done[64] = {};
  using L = H2Lay<C, EPI>;
  if (C::RESW && (a.Cout != C::N || (a.Cin + 15) / 16 > C::MAXKG)) return WCTB_E_UNSUPPORTED;
  int rc = ensure_smem_attr(conv_h2_kernel<C, EPI>, L::SMEM_BYTES, done);
  if (rc != WCTB_OK) return rc;
  a.tiles_x = (a.W + TW - 1) / TW;
  a.tiles_y = (a.H + C::TH - 1) / C::TH;
  a.ntiles = a.tiles_x * a.tiles_y * (a.Cout / C::N);
  CUtensorMap tmap;
  rc = make_h8_tmap(&tmap, a.x, a.planes_in, a.H, a.W, C::ROWS);
  if (rc != WCTB_OK) return rc;
  const int grid = a.ntiles < wctb_num_sms() ? a.ntiles : wctb_num_sms();
  conv_h2_kernel<C, EPI><<<grid, H2_THREADS, L::SMEM_BYTES, st>>>(tmap, a);
  WCTB_RETURN_LAUNCH();
}
template <class C>
int launch_h2_epi(const H2Args& a, int epi, cudaStream_t st) {
  switch (epi) {
    case WCTB_EPI_NONE: return launch_h2<C, WCTB_EPI_NONE>(a, st);
    case WCTB_EPI_POOL2: return launch_h2<C, WCTB_EPI_POOL2>(a, st);
    case WCTB_EPI_UP2: return launch_h2<C, WCTB_EPI_UP2>(a, st);
    default: return WCTB_E_BADARG;
  }
}
int g_h2_resident = 1;   // debug A/B switch (wctb_debug_set_h2_resident): 0 = round-2a kernels (weights streamed with every stage)
inline int h2_pick_n(int Cout) {
  if (Cout == 16 || Cout == 32 || Cout == 64 || Cout == 128) return Cout;
  if (Cout > 128 && Cout % 128 == 0) return 128;
  return 0;
}

// ---------------------------------------------------------------------------------- weight packing (device-side scale)
// wscale[0]: max|w| as uint bits (atomicMax; zeroed by the caller), wscale[1]: 1/s, wscale[2]: s
__global__ void h2_maxabs_kernel(const float* __restrict__ w, long long n, unsigned* __restrict__ wscale) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(wscale, __float_as_uint(m));
}
__device__ __forceinline__ float h2_scale_from_max(float m) {
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                       // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 10 - e);          // m * s in [512, 1024)
}
// dst index = ((((nb*nkg + kg)*9 + tap)*2 + c)*2N + r)*8 + e ; r < N: hi of co = nb*N + r, r >= N: lo of co = nb*N + r - N
__global__ void pack_w_h2_kernel(const float* __restrict__ w, __half* __restrict__ dst, float* __restrict__ wscale, int Cin,
                                 int Cout, int N) {
  const float s = h2_scale_from_max(__uint_as_float(reinterpret_cast<const unsigned*>(wscale)[0]));
  const int nkg = (Cin + 15) / 16;
  const long long total = (long long)(Cout / N) * nkg * 9 * 2 * 2 * N * 8;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) { wscale[1] = 1.f / s; wscale[2] = s; }
  if (i >= total) return;
  const int e = (int)(i & 7);
  long long t = i >> 3;
  const int r = (int)(t % (2 * N)); t /= (2 * N);
  const int c = (int)(t & 1); t >>= 1;
  const int tap = (int)(t % 9); t /= 9;
  const int kg = (int)(t % nkg);
  const int nb = (int)(t / nkg);
  const int co = nb * N + (r < N ? r : r - N), ci = kg * 16 + c * 8 + e;
  float v = 0.f;
  if (ci < Cin) v = w[((long long)co * Cin + ci) * 9 + tap] * s;
  const __half hi = __float2half_rn(v);
  dst[i] = (r < N) ? hi : __float2half_rn(v - __half2float(hi));
}

// ---------------------------------------------------------------------------------- layout conversion
__global__ void nchw_to_h8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int C, long long HW) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y;
  if (i >= HW) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (c8 * 8 + e < C) ? src[(long long)(c8 * 8 + e) * HW + i] : 0.f;
  uint4 hi, lo;
  split8(v, hi, lo);
  dst[(long long)(2 * c8) * HW + i] = hi;
  dst[(long long)(2 * c8 + 1) * HW + i] = lo;
}
__global__ void h8_to_nchw_kernel(const uint4* __restrict__ src, float* __restrict__ dst, int C, long long HW) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y;
  if (i >= HW) return;
  const uint4 hi = src[(long long)(2 * c8) * HW + i], lo = src[(long long)(2 * c8 + 1) * HW + i];
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t hh = (h[e >> 1] >> (16 * (e & 1))) & 0xffffu, ll = (l[e >> 1] >> (16 * (e & 1))) & 0xffffu;
    if (c8 * 8 + e < C) dst[(long long)(c8 * 8 + e) * HW + i] = h2f(hh) + h2f(ll);
  }
}
// fp32 P4 [C/4][HW][4] -> H8 (used when a decoder is fed an fp32 feature, e.g. after wct_apply)
__global__ void p4_to_h8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int C4, long long HW) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y;
  if (i >= HW) return;
  const float4 a = src[(long long)(2 * c8) * HW + i];
  const float4 b = (2 * c8 + 1 < C4) ? src[(long long)(2 * c8 + 1) * HW + i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint4 hi, lo;
  split8(v, hi, lo);
  dst[(long long)(2 * c8) * HW + i] = hi;
  dst[(long long)(2 * c8 + 1) * HW + i] = lo;
}

// ---------------------------------------------------------------------------------- first layer (3 -> Cout), fp32 FFMA
// Same arithmetic as conv_first2_kernel (conv_fp32.cu): two pixels per thread, bias then taps 0..26 in order; the
// output goes to H8 (and optionally fp32 P4 as well, when the layer is the last of its encoder).
__global__ void __launch_bounds__(256) conv_first_h2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, uint4* __restrict__ y_h8,
                                                            float4* __restrict__ y_p4, int H, int W, int Cout) {
  extern __shared__ float4 smem4[];
  float* ws = reinterpret_cast<float*>(smem4);  // [27][Cout]
  float* bs = ws + 27 * Cout;
  for (int i = threadIdx.x + threadIdx.y * 32; i < 27 * Cout; i += 256) ws[i] = w[i];
  for (int i = threadIdx.x + threadIdx.y * 32; i < Cout; i += 256) bs[i] = bias[i];
  __syncthreads();
  const int xa = blockIdx.x * 64 + threadIdx.x, xb = xa + 32;
  const int yy = blockIdx.y * 8 + threadIdx.y;
  if (xa >= W || yy >= H) return;
  const bool hasb = xb < W;
  float ina[27], inb[27];
  const long long HW = (long long)H * W;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int gy = wctb_reflect(yy + dy - 1, H);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int gxa = wctb_reflect(xa + dx - 1, W);
      const int gxb = wctb_reflect(xb + dx - 1, W);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ina[(dy * 3 + dx) * 3 + c] = __ldg(x + c * HW + (long long)gy * W + gxa);
        inb[(dy * 3 + dx) * 3 + c] = __ldg(x + c * HW + (long long)gy * W + gxb);
      }
    }
  }
  const float4* w4 = reinterpret_cast<const float4*>(ws);
  const int C4 = Cout >> 2;
  const long long row = (long long)yy * W;
  for (int c8 = 0; c8 < (Cout >> 3); ++c8) {
    float va[8], vb[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c4 = 2 * c8 + h;
      float4 a = reinterpret_cast<const float4*>(bs)[c4];
      float4 b = a;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 wv = w4[k * C4 + c4];
        a.x = fmaf(ina[k], wv.x, a.x); b.x = fmaf(inb[k], wv.x, b.x);
        a.y = fmaf(ina[k], wv.y, a.y); b.y = fmaf(inb[k], wv.y, b.y);
        a.z = fmaf(ina[k], wv.z, a.z); b.z = fmaf(inb[k], wv.z, b.z);
        a.w = fmaf(ina[k], wv.w, a.w); b.w = fmaf(inb[k], wv.w, b.w);
      }
      va[4 * h] = wctb_relu(a.x); va[4 * h + 1] = wctb_relu(a.y); va[4 * h + 2] = wctb_relu(a.z); va[4 * h + 3] = wctb_relu(a.w);
      vb[4 * h] = wctb_relu(b.x); vb[4 * h + 1] = wctb_relu(b.y); vb[4 * h + 2] = wctb_relu(b.z); vb[4 * h + 3] = wctb_relu(b.w);
      if (y_p4) {
        float4* pr = y_p4 + (long long)c4 * HW + row;
        pr[xa] = make_float4(va[4 * h], va[4 * h + 1], va[4 * h + 2], va[4 * h + 3]);
        if (hasb) pr[xb] = make_float4(vb[4 * h], vb[4 * h + 1], vb[4 * h + 2], vb[4 * h + 3]);
      }
    }
    if (y_h8) {
      uint4 hi, lo;
      uint4* ph = y_h8 + (long long)(2 * c8) * HW + row;
      uint4* pl = ph + HW;
      split8(va, hi, lo);
      ph[xa] = hi; pl[xa] = lo;
      if (hasb) { split8(vb, hi, lo); ph[xb] = hi; pl[xb] = lo; }
    }
  }
}


// ---------------------------------------------------------------------------------- microbenchmarks (debug ABI)
// cycles per tcgen05.mma.kind::f16 (M = 128, K = 16) for a given N with `nacc` independent accumulators, operands in the
// no-swizzle plane layout the conv kernels use (contents irrelevant).
template <int N>
__global__ void __launch_bounds__(128, 1) h2_mma_rate_kernel(long long* out, int nacc, int iters, int a_off16) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
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
    constexpr uint32_t idesc = umma_idesc_f16(N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = umma_desc(a0 + k * 64 * 16 + (uint32_t)((k % 3) * a_off16) * 16u, 18432u * 2u, 128u);   // taps dx = 0, 1, 2
        const uint64_t bd = umma_desc(b0 + k * 2 * N * 16, N * 16u, 128u);
#pragma unroll 1
        for (int b = 0; b < nacc; ++b) umma_f16(tmem + (uint32_t)(b * N), ad, bd, idesc, 1u);
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
// cycles for `iters` x (tcgen05.ld 32x32b.x16 of `per_iter` different column groups + wait) issued by `nwarps` warps at once
__global__ void __launch_bounds__(256, 1) h2_ldtm_rate_kernel(long long* out, int per_iter, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tq = slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int k = 0; k < per_iter; k += 2) {
      uint32_t r0[16], r1[16];
      tmem_ld16_issue(tq + (uint32_t)((16 * k) & 511), r0);
      tmem_ld16_issue(tq + (uint32_t)((16 * (k + 1)) & 511), r1);
      tmem_ld16_wait(r0);
      tmem_ld16_wait(r1);
#pragma unroll
      for (int i = 0; i < 16; ++i) sink ^= r0[i] ^ r1[i];
    }
  }
  const long long dt = clock64() - t0;
  if (lane == 0) out[blockIdx.x * 8 + warp] = dt;
  if (sink == 0x12345678u) out[63] = sink;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

}  // namespace

// ====================================================================================== C ABI
extern "C" int wctb_h2_supported(int Cin, int Cout) { return (h2_pick_n(Cout) != 0 && Cin > 0 && (Cin % 8) == 0) ? 1 : 0; }

extern "C" long long wctb_h2_packed_halves(int Cin, int Cout) {
  if (!wctb_h2_supported(Cin, Cout)) return 0;
  return (long long)Cout * ((Cin + 15) / 16) * 9 * 2 * 2 * 8;
}

extern "C" int wctb_pack_weights_h2(const float* w, void* dst, float* wscale, int Cin, int Cout, void* stream) {
  if (!w || !dst || !wscale) return WCTB_E_BADARG;
  if (!wctb_h2_supported(Cin, Cout)) return WCTB_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = 9LL * Cin * Cout;
  WCTB_CUDA_TRY(cudaMemsetAsync(wscale, 0, 4 * sizeof(float), st));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  h2_maxabs_kernel<<<blocks, 256, 0, st>>>(w, n, reinterpret_cast<unsigned*>(wscale));
  const long long total = wctb_h2_packed_halves(Cin, Cout);
  pack_w_h2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, (__half*)dst, wscale, Cin, Cout, h2_pick_n(Cout));
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_conv3x3_h2(const void* x_h8, const void* w_packed, const float* bias, const float* wscale, void* y_h8,
                               float* y_p4, int H, int W, int Cin, int Cout, int epilogue, void* stream) {
  if (!x_h8 || !w_packed || !bias || !wscale || (!y_h8 && !y_p4) || H < 2 || W < 2) return WCTB_E_BADARG;
  if (!wctb_h2_supported(Cin, Cout)) return WCTB_E_UNSUPPORTED;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  H2Args a{(const __half*)x_h8, (const __half*)w_packed, bias, wscale, (__half*)y_h8, (float4*)y_p4, H, W, Cin, Cout, 0, 0, 0,
           2 * ((Cin + 7) / 8)};
  if (epilogue == WCTB_EPI_NCHW3) {
    if (Cout != 16 || !y_p4) return WCTB_E_BADARG;
    a.y_h8 = nullptr;
    if (g_h2_resident && (Cin + 15) / 16 <= 4) return launch_h2<H2Cfg<16, 8, 1, 1>, WCTB_EPI_NCHW3>(a, st);
    return launch_h2<H2Cfg<16, 8, 1>, WCTB_EPI_NCHW3>(a, st);
  }
  // layers whose packed weights fit next to the input stages (Cout = N <= 64, Cin <= 64) keep them resident in shared memory
  const bool resw = g_h2_resident && Cout <= 64 && (Cin + 15) / 16 <= 4;
  switch (h2_pick_n(Cout)) {
    case 16: return resw ? launch_h2_epi<H2Cfg<16, 8, 1, 1>>(a, epilogue, st) : launch_h2_epi<H2Cfg<16, 8, 1>>(a, epilogue, st);
    case 32: return resw ? launch_h2_epi<H2Cfg<32, 4, 1, 1>>(a, epilogue, st) : launch_h2_epi<H2Cfg<32, 4, 1>>(a, epilogue, st);
    case 64: return resw ? launch_h2_epi<H2Cfg<64, 2, 1, 1>>(a, epilogue, st) : launch_h2_epi<H2Cfg<64, 4, 0>>(a, epilogue, st);
    default: return launch_h2_epi<H2Cfg<128, 2, 0>>(a, epilogue, st);
  }
}

extern "C" int wctb_nchw_to_h8(const float* src, void* dst, int C, int H, int W, void* stream) {
  if (!src || !dst || C <= 0 || H <= 0 || W <= 0) return WCTB_E_BADARG;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), (C + 7) / 8);
  nchw_to_h8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (uint4*)dst, C, HW);
  WCTB_RETURN_LAUNCH();
}
extern "C" int wctb_h8_to_nchw(const void* src, float* dst, int C, int H, int W, void* stream) {
  if (!src || !dst || C <= 0 || H <= 0 || W <= 0) return WCTB_E_BADARG;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), (C + 7) / 8);
  h8_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, dst, C, HW);
  WCTB_RETURN_LAUNCH();
}
extern "C" int wctb_p4_to_h8(const float* src, void* dst, int C, int H, int W, void* stream) {
  if (!src || !dst || C <= 0 || (C & 3) || H <= 0 || W <= 0) return WCTB_E_BADARG;
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), (C + 7) / 8);
  p4_to_h8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint4*)dst, C / 4, HW);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_conv3x3_first_h2(const float* x, const float* w, const float* bias, void* y_h8, float* y_p4, int H, int W,
                                     int Cout, void* stream) {
  if (!x || !w || !bias || (!y_h8 && !y_p4) || H < 2 || W < 2 || Cout <= 0 || (Cout & 7) || Cout > 512) return WCTB_E_BADARG;
  const size_t smem = (size_t)(27 * Cout + Cout) * sizeof(float);
  static bool done[64] = {};
  if (smem > 48 * 1024) {
    int rc = ensure_smem_attr(conv_first_h2_kernel, (int)smem, done);
    if (rc != WCTB_OK) return rc;
  }
  dim3 block(32, 8), grid((W + 63) / 64, (H + 7) / 8);
  conv_first_h2_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, w, bias, (uint4*)y_h8, (float4*)y_p4, H, W, Cout);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_debug_set_h2_resident(int on) { g_h2_resident = on ? 1 : 0; return WCTB_OK; }

extern "C" int wctb_debug_mma_rate_f16_off(long long* out_cycles, int N, int nacc, int iters, int ctas, int a_off16, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 200 * 1024;
  if (!out_cycles || nacc < 1 || nacc * N > 512 || iters < 1 || ctas < 1) return WCTB_E_BADARG;
#define WCTB_MR(NN) case NN: WCTB_CUDA_TRY(cudaFuncSetAttribute(h2_mma_rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    h2_mma_rate_kernel<NN><<<ctas, 128, smem, st>>>(out_cycles, nacc, iters, a_off16); break;
  switch (N) { WCTB_MR(16) WCTB_MR(32) WCTB_MR(48) WCTB_MR(64) WCTB_MR(96) WCTB_MR(128) WCTB_MR(256) default: return WCTB_E_UNSUPPORTED; }
#undef WCTB_MR
  WCTB_RETURN_LAUNCH();
}
extern "C" int wctb_debug_mma_rate_f16(long long* out_cycles, int N, int nacc, int iters, int ctas, void* stream) {
  return wctb_debug_mma_rate_f16_off(out_cycles, N, nacc, iters, ctas, 0, stream);
}
/* out: [ctas][8] cycles per warp; nwarps in {4, 8}; per_iter even */
extern "C" int wctb_debug_ldtm_rate(long long* out_cycles, int nwarps, int per_iter, int iters, int ctas, void* stream) {
  if (!out_cycles || (nwarps != 4 && nwarps != 8) || per_iter < 2 || (per_iter & 1) || iters < 1 || ctas < 1) return WCTB_E_BADARG;
  h2_ldtm_rate_kernel<<<ctas, 32 * nwarps, 0, (cudaStream_t)stream>>>(out_cycles, per_iter, iters);
  WCTB_RETURN_LAUNCH();
}
