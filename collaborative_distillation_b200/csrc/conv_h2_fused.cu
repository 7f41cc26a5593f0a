// libwctb: fused two-convolution kernels of the h2 engine (fp16 hi/lo operand pairs on tcgen05.mma.kind::f16).
//
//   head:  y = pool2( ReLU(conv12( pad( ReLU(conv11( pad(img) )) ) )) )       img NCHW fp32 3ch  -> y H8 16ch, half res
//   tail:  img = ReLU(conv11( pad( ReLU(conv12( pad( [up2] x ) )) ) ))       x H8 16ch           -> img NCHW fp32 3ch
//
// The 16-channel full-resolution layers of the 16x nets are bound by the A-operand fetch of tcgen05.mma: one 128-row
// operand slab (4 KB) costs ~40 cycles whatever N is (profiles/r02_h2_rates.txt: N=16 39, N=48 44, N=96 56 cycles).
// So here the three HORIZONTAL taps of a 3x3 filter are stacked along N instead of being three MMAs:
//      D[p, dx*16 + co] = sum_{dy, ci} A[p + dy*pitch, ci] * W[co, ci, dy, dx]            (one MMA per dy, N = 3*16 [*2])
//      out[p, co]       = D[p, 0*16+co] + D[p+1, 1*16+co] + D[p+2, 2*16+co]
// and the shifted sum is done by the epilogue with two warp shuffles per channel.  That needs p, p+1, p+2 in one warp:
// the tile pitch is 32 pixels -- a TMEM lane quarter (= one warp) is exactly one tile row, the two rightmost positions
// of a row are the usual garbage columns, an accumulator block of 128 positions is 4 rows.  With the hi/lo weight halves
// stacked as well (N = 96: [3 dx][16] main | [3 dx][16] minor) a 16->16 conv costs 3 x (56 + 44) = 300 cycles per 128
// positions instead of 9 x (40 + 39) = 711.
//
// Pipeline.  A tile is 32 output rows x 28 columns.  Inside a tile the two convolutions are pipelined at accumulator-
// block granularity through two small TMEM rings (2 x 96 columns each): the first conv's block j is converted by its
// epilogue warps into rows 4j..4j+3 of the second conv's operand tile in shared memory (hi/lo split, reflection of the
// INTERMEDIATE at true image borders), the second conv's block k starts as soon as rows 4k..4k+5 are there, and its
// epilogue warps pool / store while the single MMA-issuing thread is already ahead.  Image and operand tiles are double
// buffered, so the pipeline also runs across tile boundaries; CTAs are persistent (one per SM).
//   warp 0      weight loader (bulk copies, once)           warp 1      tcgen05.mma issuer (one elected thread)
//   warps 2-5   epilogue of the second conv                 warps 6-9   epilogue of the first conv -> operand tile
//   warps 10-13 first-conv input: image loader / converter (head) or upsampling loader (tail, when the input is half res)
#include "h2.cuh"

namespace {
using namespace wctb_umma;

constexpr int FP = 32;                              // tile pitch (pixels) = warp width
constexpr int F_TW = 28;                            // valid output columns after two chained 3x3 convs (32 -> 30 -> 28)
constexpr int F_ROW = FP * 16;                      // bytes of one 8-channel row
constexpr int F_WB = 2 * 96 * 16;                   // one B tile: [2 k-chunks][96 rows][16 B]
constexpr int F_ACC1 = 0, F_ACC2 = 192;             // TMEM column bases of the two rings (2 slots x 96 columns each)
// warps: 0 weights, 1 MMA issuer, 2-9 second-conv epilogue (two groups of 4), 10-17 first-conv epilogue (two groups of 4),
// 18-21 loader.  The two groups of an epilogue alternate accumulator blocks (group = TMEM ring slot = block parity), so a
// group has two block periods for its ~300-instruction body and the SM sub-partitions always have a second warp to issue from.
constexpr int F_THREADS = 22 * 32;
// barrier slots (B_MID_READY is followed by 2 x NB1 barriers)
enum { B_WFULL = 0, B_IN_READY = 1, B_IN_FREE = 3, B_A1_FULL = 5, B_A1_EMPTY = 7, B_A2_FULL = 9, B_A2_EMPTY = 11, B_MID_FREE = 13,
       B_MID_READY = 15 };

template <int NB2_>
struct FCfg {
  static constexpr int NB2 = NB2_;                  // second-conv blocks (4 rows x 32 positions) per full tile
  static constexpr int NB1 = NB2_ + 1;              // first-conv blocks per full tile (tile rows -1 .. 4*NB2 + 2)
  static constexpr int TH = 4 * NB2_;               // output rows per tile
  static constexpr int MID_ROWS = 4 * NB1;
  static constexpr int PLANE = MID_ROWS * F_ROW;
  static constexpr int MID_BYTES = 4 * PLANE;       // planes hi0, lo0, hi1, lo1
  static constexpr int NBAR = B_MID_READY + 2 * NB1;
};

struct FTile { int x0, ya, nb2, nb1; };
template <class C>
__device__ __forceinline__ FTile f_tile(int tile, int tiles_x, int H) {
  FTile t;
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  t.x0 = tx * F_TW;
  t.ya = ty * C::TH;
  const int rows = min(C::TH, H - t.ya);
  t.nb2 = (rows + 3) >> 2;
  t.nb1 = t.nb2 + 1;
  return t;
}

// out[c] = sum_dx shfl_down(main[dx*16 + c] + minor[dx*16 + c], dx)   for one dx-stacked accumulator block (96 columns)
__device__ __forceinline__ void f_reduce_dx(uint32_t taddr, float* out) {
  uint32_t m0[16], n0[16], m1[16], n1[16];
  tmem_ld16_issue(taddr, m0);
  tmem_ld16_issue(taddr + 48u, n0);
  tmem_ld16_issue(taddr + 16u, m1);
  tmem_ld16_issue(taddr + 64u, n1);
  tmem_ld16_wait(m0);
  tmem_ld16_wait(n0);
  tmem_ld16_wait(m1);
  tmem_ld16_wait(n1);
  f32x2 acc[8];
  float s1[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = add2(pk2(__uint_as_float(m0[2 * i]), __uint_as_float(m0[2 * i + 1])), pk2(__uint_as_float(n0[2 * i]), __uint_as_float(n0[2 * i + 1])));
    upk2(add2(pk2(__uint_as_float(m1[2 * i]), __uint_as_float(m1[2 * i + 1])), pk2(__uint_as_float(n1[2 * i]), __uint_as_float(n1[2 * i + 1]))),
         s1[2 * i], s1[2 * i + 1]);
  }
  tmem_ld16_issue(taddr + 32u, m0);              // dx = 2 flies while the dx = 1 shuffles run
  tmem_ld16_issue(taddr + 80u, n0);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    acc[i] = add2(acc[i], pk2(__shfl_down_sync(0xffffffffu, s1[2 * i], 1), __shfl_down_sync(0xffffffffu, s1[2 * i + 1], 1)));
  tmem_ld16_wait(m0);
  tmem_ld16_wait(n0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, b;
    upk2(add2(pk2(__uint_as_float(m0[2 * i]), __uint_as_float(m0[2 * i + 1])), pk2(__uint_as_float(n0[2 * i]), __uint_as_float(n0[2 * i + 1]))), a, b);
    acc[i] = add2(acc[i], pk2(__shfl_down_sync(0xffffffffu, a, 2), __shfl_down_sync(0xffffffffu, b, 2)));
    upk2(acc[i], out[2 * i], out[2 * i + 1]);
  }
}
// v[i] = relu(v[i] * s + b[i]) for 16 values, packed FFMA2
__device__ __forceinline__ void f_scale_bias_relu16(float* v, float s, const float* b) {
  const f32x2 s2 = pk2(s, s);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, c;
    upk2(fma2(pk2(v[2 * i], v[2 * i + 1]), s2, pk2(b[2 * i], b[2 * i + 1])), a, c);
    v[2 * i] = wctb_relu(a);
    v[2 * i + 1] = wctb_relu(c);
  }
}

// second conv of a chain (16 -> 16 or 16 -> 3 padded), dx-stacked: block k of the operand tile `mid` ([hi0, lo0, hi1, lo1]
// planes, pitch 32) -> TMEM columns [tacc, tacc + 96).  wsm: [3 dy][2 chunks][96 rows][16 B] (rows 0..47 hi, 48..95 lo).
template <int PLANE>
__device__ __forceinline__ void f_issue_conv16(uint32_t mid, uint32_t wsm, uint32_t tacc, int k) {
  constexpr uint32_t id96 = umma_idesc_f16(96), id48 = umma_idesc_f16(48);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const uint32_t a = mid + (uint32_t)(4 * k + dy) * F_ROW;
    const uint64_t bd = umma_desc(wsm + (uint32_t)dy * F_WB, 96u * 16u, 128u);
    umma_f16(tacc, umma_desc(a, 2u * PLANE, 128u), bd, id96, dy > 0 ? 1u : 0u);            // hi x [w_hi | w_lo]
    umma_f16(tacc, umma_desc(a + PLANE, 2u * PLANE, 128u), bd, id48, 1u);                  // lo x  w_hi
  }
}

// write 16 channels of one position into the operand tile (planes hi0, lo0, hi1, lo1)
template <int PLANE>
__device__ __forceinline__ void f_store_mid(uint8_t* mid, int row, int col, const float* v) {
  uint4 hi, lo;
  const uint32_t p = smem_u32(mid) + (uint32_t)(row * F_ROW + col * 16);
  split8(v, hi, lo);
  sts128(p, hi);
  sts128(p + PLANE, lo);
  split8(v + 8, hi, lo);
  sts128(p + 2 * PLANE, hi);
  sts128(p + 3 * PLANE, lo);
}
template <int PLANE>
__device__ __forceinline__ void f_copy_mid_row(uint8_t* mid, int dst_row, int src_row, int lane) {
#pragma unroll
  for (int pl = 0; pl < 4; ++pl) {
    const uint32_t base = smem_u32(mid) + (uint32_t)(pl * PLANE + lane * 16);
    sts128(base + dst_row * F_ROW, lds128(base + src_row * F_ROW));
  }
}

// epilogue of the FIRST conv of a chain: accumulator block j -> rows 4j..4j+3 of the operand tile, with the reflection
// of the intermediate at true image borders (columns by shuffle before the store, rows by a copy after a group barrier)
template <class C>
__device__ __forceinline__ void f_first_epilogue(uint32_t tacc_q, uint8_t* mid, const FTile& t, int j, int q, int lane,
                                                 int H, int W, const float* bias, float inv_s, uint64_t* acc_empty, int grp,
                                                 uint64_t* prev_ready, uint32_t ready_parity) {
  float v[16];
  f_reduce_dx(tacc_q, v);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(acc_empty);
  f_scale_bias_relu16(v, inv_s, bias);
  // tile col c <-> gx = x0 - 1 + c.  gx = -1 takes the value of gx = 1, gx = W the value of gx = W - 2
  const bool left = (t.x0 == 0);
  const int cr = W - t.x0 + 1;
  const bool right = cr < 30;
  if (left || right) {
    int src = lane;
    if (left && lane == 0) src = 2;
    if (right && lane == cr) src = cr - 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], src);
  }
  const int row = 4 * j + q;
  f_store_mid<C::PLANE>(mid, row, lane, v);
  // tile row r <-> gy = ya - 1 + r.  gy = -1 takes row gy = 1, gy = H takes row gy = H - 2
  const bool top = (t.ya == 0) && (j == 0);
  const int rh = H - t.ya + 1;
  const bool bottom = (H - t.ya <= C::TH) && (j == (rh >> 2));
  if (top || bottom) {
    // the two epilogue groups alternate blocks: rows of this block were written by this group (named barrier), the source
    // row of the bottom patch may sit in the previous block, written by the other group (its mid_ready barrier)
    if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
    else asm volatile("bar.sync 4, 128;" ::: "memory");
    if (top && q == 0) f_copy_mid_row<C::PLANE>(mid, 0, 2, lane);
    if (bottom && q == (rh & 3)) {
      if (j > 0) mbar_wait_sleep(prev_ready, ready_parity);
      f_copy_mid_row<C::PLANE>(mid, rh, rh - 2, lane);
    }
  }
}

// ====================================================================================== fused encoder head
struct HeadH2Args {
  const float* img;     // [3][H][W]
  const __half* w11;    // [2 mma][2 chunks][96][8]   (conv0 folded; see ops.pack_head_h2_w11)
  const __half* w12;    // [3 dy][2 chunks][96][8]
  const float* b11;     // [16]
  const float* b12;     // [16]
  float inv_s11, inv_s12;
  uint4* y;             // H8 [2 chunks][2][H/2][W/2] 16-byte units
  int H, W, tiles_x, ntiles;
};
using HC = FCfg<8>;                                           // head tile: 32 rows x 28 columns
constexpr int HD_IN_ROWS = 40;                                // 36 image rows + overrun of the last block's second K chunk
constexpr int HD_IN_BYTES = HD_IN_ROWS * F_ROW;               // RGB0 hi | RGB0 lo, 16 B per pixel
constexpr int HD_OFF_MID = 2 * HD_IN_BYTES;
constexpr int HD_OFF_W11 = HD_OFF_MID + 2 * HC::MID_BYTES;
constexpr int HD_OFF_W12 = HD_OFF_W11 + 2 * F_WB;
constexpr int HD_OFF_POOL = HD_OFF_W12 + 3 * F_WB;
constexpr int HD_POOL_BYTES = 2 * 2 * 2 * 32 * 20 * 4;        // [parity][group][row pair][lane][16 + 4 pad] floats
constexpr int HD_OFF_BAR = HD_OFF_POOL + HD_POOL_BYTES;
constexpr int HD_SMEM = HD_OFF_BAR + 512 + 128;
static_assert(HC::NBAR * 8 + 8 <= 512, "barrier area");
static_assert(HD_SMEM <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(F_THREADS, 1) conv_head_h2_kernel(const HeadH2Args h) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HD_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + HC::NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = h.H, W = h.W;
  const int ntl = (h.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // tiles of this CTA

  if (threadIdx.x == 0) {
    mbar_init(bars + B_WFULL, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + B_IN_READY + s, 4);   mbar_init(bars + B_IN_FREE + s, 1);
      mbar_init(bars + B_A1_FULL + s, 1);    mbar_init(bars + B_A1_EMPTY + s, 4);
      mbar_init(bars + B_A2_FULL + s, 1);    mbar_init(bars + B_A2_EMPTY + s, 4);
      mbar_init(bars + B_MID_FREE + s, 1);
      for (int j = 0; j < HC::NB1; ++j) mbar_init(bars + B_MID_READY + s * HC::NB1 + j, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== weights (once) ===========================
    if (elect_one()) {
      mbar_expect_tx(bars + B_WFULL, 5u * F_WB);
      bulk_g2s(smem_u32(smem + HD_OFF_W11), h.w11, 2u * F_WB, bars + B_WFULL);
      bulk_g2s(smem_u32(smem + HD_OFF_W12), h.w12, 3u * F_WB, bars + B_WFULL);
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t id96 = umma_idesc_f16(96);
      const uint32_t w11s = smem_u32(smem + HD_OFF_W11), w12s = smem_u32(smem + HD_OFF_W12);
      mbar_wait_sleep(bars + B_WFULL, 0);
      int i1 = 0, j1 = 0, i2 = 0, k2 = 0;          // cursors: (local tile, block) of the next first- / second-conv block
      uint32_t g1 = 0, g2 = 0, base1 = 0;          // global block counters; first-conv blocks issued before tile i2
      FTile t1 = f_tile<HC>((int)blockIdx.x, h.tiles_x, H), t2 = t1;
      while (i2 < ntl) {
        // second-conv block (i2, k2) reads first-conv blocks 0..k2+1 of its tile; stay one more block ahead
        const uint32_t need = base1 + (uint32_t)k2 + 2u;
        while (i1 < ntl && g1 < need + 1u) {
          if (j1 == 0) mbar_wait_sleep(bars + B_IN_READY + (i1 & 1), (i1 >> 1) & 1);
          const uint32_t slot = g1 & 1u;
          mbar_wait_sleep(bars + B_A1_EMPTY + slot, ((g1 >> 1) & 1u) ^ 1u);
          tc_fence_after();
          // conv11: K chunk 0 = image row r (RGB hi | RGB lo of one pixel), chunk 1 = row r + 1 (LBO = one row)
          const uint32_t a = smem_u32(smem + (i1 & 1) * HD_IN_BYTES) + (uint32_t)(4 * j1) * F_ROW;
          const uint32_t tacc = tmem_base + F_ACC1 + slot * 96u;
          umma_f16(tacc, umma_desc(a, F_ROW, 128u), umma_desc(w11s, 96u * 16u, 128u), id96, 0u);                  // dy 0, 1
          umma_f16(tacc, umma_desc(a + 2u * F_ROW, F_ROW, 128u), umma_desc(w11s + F_WB, 96u * 16u, 128u), id96, 1u);  // dy 2, (zero)
          tc_commit(bars + B_A1_FULL + slot);
          ++g1;
          if (++j1 == t1.nb1) {
            tc_commit(bars + B_IN_FREE + (i1 & 1));
            j1 = 0;
            ++i1;
            if (i1 < ntl) t1 = f_tile<HC>((int)blockIdx.x + i1 * (int)gridDim.x, h.tiles_x, H);
          }
        }
        {
          uint64_t* ready = bars + B_MID_READY + (i2 & 1) * HC::NB1;
          mbar_wait_sleep(ready + k2, (i2 >> 1) & 1);
          mbar_wait_sleep(ready + k2 + 1, (i2 >> 1) & 1);
          const uint32_t slot = g2 & 1u;
          mbar_wait_sleep(bars + B_A2_EMPTY + slot, ((g2 >> 1) & 1u) ^ 1u);
          tc_fence_after();
          f_issue_conv16<HC::PLANE>(smem_u32(smem + HD_OFF_MID + (i2 & 1) * HC::MID_BYTES), w12s, tmem_base + F_ACC2 + slot * 96u, k2);
          tc_commit(bars + B_A2_FULL + slot);
          ++g2;
          if (++k2 == t2.nb2) {
            tc_commit(bars + B_MID_FREE + (i2 & 1));
            base1 += (uint32_t)t2.nb1;
            k2 = 0;
            ++i2;
            if (i2 < ntl) t2 = f_tile<HC>((int)blockIdx.x + i2 * (int)gridDim.x, h.tiles_x, H);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // =========================== epilogue of conv12: pool, split, store ===========================
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t tq = tmem_base + F_ACC2 + ((uint32_t)(32 * q) << 16);
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(h.b12 + i);
    float* poolbuf = reinterpret_cast<float*>(smem + HD_OFF_POOL);
    const int Ho = H >> 1, Wo = W >> 1;
    const long long HWo = (long long)Ho * Wo;
    uint32_t g2 = 0;
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<HC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      for (int k = 0; k < t.nb2; ++k, ++g2) {
        const uint32_t slot = g2 & 1u;
        if (slot != (uint32_t)grp) continue;        // the other group's block
        mbar_wait_sleep(bars + B_A2_FULL + slot, (g2 >> 1) & 1u);
        tc_fence_after();
        float v[16];
        f_reduce_dx(tq + slot * 96u, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_A2_EMPTY + slot);
        f_scale_bias_relu16(v, h.inv_s12, bv);
        // rows 4k+q: (q, q+1) pool together; odd rows hand their values to the even row's warp
        float* pb = poolbuf + ((((g2 >> 1) & 1u) * 2u + (uint32_t)grp) * 2u + (uint32_t)(q >> 1)) * (32 * 20) + lane * 20;
        if (q & 1) {
          float4* d = reinterpret_cast<float4*>(pb);
          d[0] = make_float4(v[0], v[1], v[2], v[3]); d[1] = make_float4(v[4], v[5], v[6], v[7]);
          d[2] = make_float4(v[8], v[9], v[10], v[11]); d[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (!(q & 1)) {
          const float4* s = reinterpret_cast<const float4*>(pb);
          const float4 s0 = s[0], s1 = s[1], s2 = s[2], s3 = s[3];
          const float o[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float m = fmaxf(v[c], o[c]);
            v[c] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          }
          const int oy = (t.ya + 4 * k + q) >> 1, ox = (t.x0 + lane) >> 1;
          if (!(lane & 1) && lane < F_TW && oy < Ho && ox < Wo) {
            const long long off = (long long)oy * Wo + ox;
            uint4 hi, lo;
            split8(v, hi, lo);
            h.y[off] = hi;
            h.y[HWo + off] = lo;
            split8(v + 8, hi, lo);
            h.y[2 * HWo + off] = hi;
            h.y[3 * HWo + off] = lo;
          }
        }
      }
    }
  } else if (warp < 18) {
    // =========================== epilogue of conv11 -> conv12 operand tile ===========================
    const int q = warp & 3, grp = (warp - 10) >> 2;
    const uint32_t tq = tmem_base + F_ACC1 + ((uint32_t)(32 * q) << 16);
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(h.b11 + i);
    uint32_t g1 = 0;
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<HC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      uint8_t* mid = smem + HD_OFF_MID + (i & 1) * HC::MID_BYTES;
      if (i >= 2) mbar_wait_sleep(bars + B_MID_FREE + (i & 1), ((i - 2) >> 1) & 1);      // conv12 of tile i-2 has read this buffer
      for (int j = 0; j < t.nb1; ++j, ++g1) {
        const uint32_t slot = g1 & 1u;
        if (slot != (uint32_t)grp) continue;        // the other group's block
        mbar_wait_sleep(bars + B_A1_FULL + slot, (g1 >> 1) & 1u);
        tc_fence_after();
        f_first_epilogue<HC>(tq + slot * 96u, mid, t, j, q, lane, H, W, bv, h.inv_s11, bars + B_A1_EMPTY + slot, grp,
                             bars + B_MID_READY + (i & 1) * HC::NB1 + (j > 0 ? j - 1 : 0), (uint32_t)((i >> 1) & 1));
        fence_async_smem();                         // every writer: generic-proxy stores -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_MID_READY + (i & 1) * HC::NB1 + j);
      }
    }
  } else {
    // =========================== image loader: NCHW fp32 -> [R G B 0 | r g b 0] fp16 hi | lo pixels ===========================
    const int pw = warp - 18;
    const long long HW = (long long)H * W;
    for (int b = 0; b < 2; ++b)      // rows 36..39 are only read against zero weights / by garbage positions: keep them finite
      reinterpret_cast<uint4*>(smem + b * HD_IN_BYTES)[(36 + pw) * FP + lane] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<HC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      uint4* in = reinterpret_cast<uint4*>(smem + (i & 1) * HD_IN_BYTES);
      if (i >= 2) mbar_wait_sleep(bars + B_IN_FREE + (i & 1), ((i - 2) >> 1) & 1);
      const int gx = wctb_reflect(t.x0 - 2 + lane, W);
      float r0[9], r1[9], r2[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {                                  // tile row r <-> gy = ya - 2 + r
        const int gy = wctb_reflect(t.ya - 2 + pw + 4 * k, H);
        const float* p = h.img + (long long)gy * W + gx;
        r0[k] = __ldg(p); r1[k] = __ldg(p + HW); r2[k] = __ldg(p + 2 * HW);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const uint32_t h0 = f2h_sat(r0[k]), h1 = f2h_sat(r1[k]), h2 = f2h_sat(r2[k]);
        const uint32_t l0 = f2h_sat(r0[k] - h2f(h0)), l1 = f2h_sat(r1[k] - h2f(h1)), l2 = f2h_sat(r2[k] - h2f(h2));
        sts128(smem_u32(in + (pw + 4 * k) * FP + lane), make_uint4(h0 | (h1 << 16), h2, l0 | (l1 << 16), l2));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_IN_READY + (i & 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ====================================================================================== fused decoder tail
struct TailH2Args {
  const uint4* x;       // H8 16 ch: [4 planes hi0, lo0, hi1, lo1][Hs][Ws] 16-byte units; (Hs, Ws) = (H, W) >> ups
  const __half* w12;    // [3 dy][2 chunks][96][8]
  const __half* w11;    // [3 dy][2 chunks][96][8], output channels 3..15 zero
  const float* b12;     // [16]
  const float* b11;     // [3]
  float inv_s12, inv_s11;
  float* img;           // [3][H][W]  (sharded: the rank's NEXT-stage extended strip, see TailShard)
  int H, W, ups, tiles_x, ntiles;
  // Strip-sharded output (multi-GPU): only the rank's own columns [own_x0, own_x0 + own_w) of the computed image are kept.
  // They go to the local buffer at column out_x0 (pitch out_pitch) and, for the `halo` columns next to a seam, ALSO straight
  // into the neighbour's next-stage buffer through its peer-mapped pointer (st.global over NVLink): the halo exchange of
  // the next stage happens inside this kernel's epilogue instead of a pack / send / recv / unpack sequence.
  // Unsharded: own_x0 = 0, own_w = W, out_pitch = W, out_x0 = 0, peers null.
  int own_x0, own_w, out_pitch, out_x0, halo;
  long long out_plane;
  float* peer_l; int peer_l_pitch, peer_l_x0; long long peer_l_plane;     // our first `halo` own columns -> left neighbour's right halo
  float* peer_r; int peer_r_pitch, peer_r_x0; long long peer_r_plane;     // our last `halo` own columns -> right neighbour's left halo
};
using TC = FCfg<4>;                                           // tail tile: 16 rows x 28 columns
constexpr int TL_IN_ROWS = 4 * TC::NB1 + 2;                   // 22 input rows (tile rows -2 .. 19)
constexpr int TL_IN_PLANE = TL_IN_ROWS * F_ROW;
constexpr int TL_IN_BYTES = 4 * TL_IN_PLANE;
constexpr int TL_OFF_MID = 2 * TL_IN_BYTES;
constexpr int TL_OFF_W12 = TL_OFF_MID + 2 * TC::MID_BYTES;
constexpr int TL_OFF_W11 = TL_OFF_W12 + 3 * F_WB;
constexpr int TL_OFF_BAR = TL_OFF_W11 + 3 * F_WB;
constexpr int TL_SMEM = TL_OFF_BAR + 512 + 128;
static_assert(TC::NBAR * 8 + 8 <= 512, "barrier area");
static_assert(TL_SMEM <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(F_THREADS, 1) conv_tail_h2_kernel(const TailH2Args h) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TL_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC::NBAR);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = h.H, W = h.W;
  const int ntl = (h.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(bars + B_WFULL, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bars + B_IN_READY + s, 4);   mbar_init(bars + B_IN_FREE + s, 1);
      mbar_init(bars + B_A1_FULL + s, 1);    mbar_init(bars + B_A1_EMPTY + s, 4);
      mbar_init(bars + B_A2_FULL + s, 1);    mbar_init(bars + B_A2_EMPTY + s, 4);
      mbar_init(bars + B_MID_FREE + s, 1);
      for (int j = 0; j < TC::NB1; ++j) mbar_init(bars + B_MID_READY + s * TC::NB1 + j, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bars + B_WFULL, 6u * F_WB);
      bulk_g2s(smem_u32(smem + TL_OFF_W12), h.w12, 3u * F_WB, bars + B_WFULL);
      bulk_g2s(smem_u32(smem + TL_OFF_W11), h.w11, 3u * F_WB, bars + B_WFULL);
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (same schedule as the head) ===========================
    if (elect_one()) {
      const uint32_t w12s = smem_u32(smem + TL_OFF_W12), w11s = smem_u32(smem + TL_OFF_W11);
      mbar_wait_sleep(bars + B_WFULL, 0);
      int i1 = 0, j1 = 0, i2 = 0, k2 = 0;
      uint32_t g1 = 0, g2 = 0, base1 = 0;
      FTile t1 = f_tile<TC>((int)blockIdx.x, h.tiles_x, H), t2 = t1;
      while (i2 < ntl) {
        const uint32_t need = base1 + (uint32_t)k2 + 2u;
        while (i1 < ntl && g1 < need + 1u) {
          if (j1 == 0) mbar_wait_sleep(bars + B_IN_READY + (i1 & 1), (i1 >> 1) & 1);
          const uint32_t slot = g1 & 1u;
          mbar_wait_sleep(bars + B_A1_EMPTY + slot, ((g1 >> 1) & 1u) ^ 1u);
          tc_fence_after();
          f_issue_conv16<TL_IN_PLANE>(smem_u32(smem + (i1 & 1) * TL_IN_BYTES), w12s, tmem_base + F_ACC1 + slot * 96u, j1);
          tc_commit(bars + B_A1_FULL + slot);
          ++g1;
          if (++j1 == t1.nb1) {
            tc_commit(bars + B_IN_FREE + (i1 & 1));
            j1 = 0;
            ++i1;
            if (i1 < ntl) t1 = f_tile<TC>((int)blockIdx.x + i1 * (int)gridDim.x, h.tiles_x, H);
          }
        }
        {
          uint64_t* ready = bars + B_MID_READY + (i2 & 1) * TC::NB1;
          mbar_wait_sleep(ready + k2, (i2 >> 1) & 1);
          mbar_wait_sleep(ready + k2 + 1, (i2 >> 1) & 1);
          const uint32_t slot = g2 & 1u;
          mbar_wait_sleep(bars + B_A2_EMPTY + slot, ((g2 >> 1) & 1u) ^ 1u);
          tc_fence_after();
          f_issue_conv16<TC::PLANE>(smem_u32(smem + TL_OFF_MID + (i2 & 1) * TC::MID_BYTES), w11s, tmem_base + F_ACC2 + slot * 96u, k2);
          tc_commit(bars + B_A2_FULL + slot);
          ++g2;
          if (++k2 == t2.nb2) {
            tc_commit(bars + B_MID_FREE + (i2 & 1));
            base1 += (uint32_t)t2.nb1;
            k2 = 0;
            ++i2;
            if (i2 < ntl) t2 = f_tile<TC>((int)blockIdx.x + i2 * (int)gridDim.x, h.tiles_x, H);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 10) {
    // =========================== epilogue of conv11: 3 channels -> NCHW fp32 ===========================
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const uint32_t tq = tmem_base + F_ACC2 + ((uint32_t)(32 * q) << 16);
    const float b0 = __ldg(h.b11), b1 = __ldg(h.b11 + 1), b2 = __ldg(h.b11 + 2);
    uint32_t g2 = 0;
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<TC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      for (int k = 0; k < t.nb2; ++k, ++g2) {
        const uint32_t slot = g2 & 1u;
        if (slot != (uint32_t)grp) continue;        // the other group's block
        mbar_wait_sleep(bars + B_A2_FULL + slot, (g2 >> 1) & 1u);
        tc_fence_after();
        // channels 0..2 of each dx group: main at dx*16, minor at 48 + dx*16
        float m[6][4];
        const uint32_t ta = tq + slot * 96u;
#pragma unroll
        for (int d = 0; d < 3; ++d) { tmem_ld4(ta + 16u * d, m[d]); tmem_ld4(ta + 48u + 16u * d, m[3 + d]); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_A2_EMPTY + slot);
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          v[c] = (m[0][c] + m[3][c]) + __shfl_down_sync(0xffffffffu, m[1][c] + m[4][c], 1) + __shfl_down_sync(0xffffffffu, m[2][c] + m[5][c], 2);
        const int gy = t.ya + 4 * k + q, gx = t.x0 + lane;
        const int ox = gx - h.own_x0;
        if (lane < F_TW && gy < H && gx < W && ox >= 0 && ox < h.own_w) {
          const float r0 = wctb_relu(fmaf(v[0], h.inv_s11, b0)), r1 = wctb_relu(fmaf(v[1], h.inv_s11, b1)),
                      r2 = wctb_relu(fmaf(v[2], h.inv_s11, b2));
          const long long o = (long long)gy * h.out_pitch + ox + h.out_x0;
          h.img[o] = r0;
          h.img[h.out_plane + o] = r1;
          h.img[2 * h.out_plane + o] = r2;
          if (h.peer_l && ox < h.halo) {                      // peer store: the left neighbour's right halo
            const long long p = (long long)gy * h.peer_l_pitch + ox + h.peer_l_x0;
            h.peer_l[p] = r0; h.peer_l[h.peer_l_plane + p] = r1; h.peer_l[2 * h.peer_l_plane + p] = r2;
          }
          if (h.peer_r && ox >= h.own_w - h.halo) {           // peer store: the right neighbour's left halo
            const long long p = (long long)gy * h.peer_r_pitch + (ox - (h.own_w - h.halo)) + h.peer_r_x0;
            h.peer_r[p] = r0; h.peer_r[h.peer_r_plane + p] = r1; h.peer_r[2 * h.peer_r_plane + p] = r2;
          }
        }
      }
    }
  } else if (warp < 18) {
    // =========================== epilogue of conv12 -> conv11 operand tile ===========================
    const int q = warp & 3, grp = (warp - 10) >> 2;
    const uint32_t tq = tmem_base + F_ACC1 + ((uint32_t)(32 * q) << 16);
    float bv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bv[i] = __ldg(h.b12 + i);
    uint32_t g1 = 0;
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<TC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      uint8_t* mid = smem + TL_OFF_MID + (i & 1) * TC::MID_BYTES;
      if (i >= 2) mbar_wait_sleep(bars + B_MID_FREE + (i & 1), ((i - 2) >> 1) & 1);
      for (int j = 0; j < t.nb1; ++j, ++g1) {
        const uint32_t slot = g1 & 1u;
        if (slot != (uint32_t)grp) continue;        // the other group's block
        mbar_wait_sleep(bars + B_A1_FULL + slot, (g1 >> 1) & 1u);
        tc_fence_after();
        f_first_epilogue<TC>(tq + slot * 96u, mid, t, j, q, lane, H, W, bv, h.inv_s12, bars + B_A1_EMPTY + slot, grp,
                             bars + B_MID_READY + (i & 1) * TC::NB1 + (j > 0 ? j - 1 : 0), (uint32_t)((i >> 1) & 1));
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + B_MID_READY + (i & 1) * TC::NB1 + j);
      }
    }
  } else {
    // =========================== input loader: (up-sampled) H8 tile, reflection resolved in the source address ===========================
    const int pw = warp - 18;
    const int Hs = H >> h.ups, Ws = W >> h.ups;
    const long long HWs = (long long)Hs * Ws;
    for (int i = 0; i < ntl; ++i) {
      const FTile t = f_tile<TC>((int)blockIdx.x + i * (int)gridDim.x, h.tiles_x, H);
      uint4* in = reinterpret_cast<uint4*>(smem + (i & 1) * TL_IN_BYTES);
      if (i >= 2) mbar_wait_sleep(bars + B_IN_FREE + (i & 1), ((i - 2) >> 1) & 1);
      const int sx = wctb_reflect(t.x0 - 2 + lane, W) >> h.ups;               // tile col c <-> gx = x0 - 2 + c
      // warp pw loads plane pw: 22 rows, two batches of 11 loads in flight
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 v[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const int r = half * 11 + k;                                        // tile row r <-> gy = ya - 2 + r
          const int sy = wctb_reflect(t.ya - 2 + r, H) >> h.ups;
          v[k] = __ldg(h.x + (long long)pw * HWs + (long long)sy * Ws + sx);
        }
#pragma unroll
        for (int k = 0; k < 11; ++k) sts128(smem_u32(in + (pw * TL_IN_ROWS + half * 11 + k) * FP + lane), v[k]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + B_IN_READY + (i & 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace

// ====================================================================================== C ABI
extern "C" int wctb_conv_head_h2(const float* x_nchw, const void* w11_packed, const float* b11, float inv_s11,
                                 const void* w12_packed, const float* b12, float inv_s12, void* y_h8, int H, int W,
                                 void* stream) {
  if (!x_nchw || !w11_packed || !b11 || !w12_packed || !b12 || !y_h8 || H < 2 || W < 2) return WCTB_E_BADARG;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  static bool done[64] = {};
  int rc = ensure_smem_attr(conv_head_h2_kernel, HD_SMEM, done);
  if (rc != WCTB_OK) return rc;
  HeadH2Args h{x_nchw, (const __half*)w11_packed, (const __half*)w12_packed, b11, b12, inv_s11, inv_s12, (uint4*)y_h8, H, W, 0, 0};
  h.tiles_x = (W + F_TW - 1) / F_TW;
  h.ntiles = h.tiles_x * ((H + HC::TH - 1) / HC::TH);
  const int grid = h.ntiles < wctb_num_sms() ? h.ntiles : wctb_num_sms();
  conv_head_h2_kernel<<<grid, F_THREADS, HD_SMEM, (cudaStream_t)stream>>>(h);
  WCTB_RETURN_LAUNCH();
}

static int launch_tail_h2(TailH2Args h, cudaStream_t st) {
  static bool done[64] = {};
  int rc = ensure_smem_attr(conv_tail_h2_kernel, TL_SMEM, done);
  if (rc != WCTB_OK) return rc;
  h.tiles_x = (h.W + F_TW - 1) / F_TW;
  h.ntiles = h.tiles_x * ((h.H + TC::TH - 1) / TC::TH);
  const int grid = h.ntiles < wctb_num_sms() ? h.ntiles : wctb_num_sms();
  conv_tail_h2_kernel<<<grid, F_THREADS, TL_SMEM, st>>>(h);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_conv_tail_h2(const void* x_h8, const void* w12_packed, const float* b12, float inv_s12,
                                 const void* w11_packed, const float* b11, float inv_s11, float* y_nchw, int H, int W,
                                 int upsample_input, void* stream) {
  if (!x_h8 || !w12_packed || !b12 || !w11_packed || !b11 || !y_nchw || H < 2 || W < 2) return WCTB_E_BADARG;
  if (upsample_input && ((H & 1) || (W & 1))) return WCTB_E_BADARG;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  TailH2Args h{(const uint4*)x_h8, (const __half*)w12_packed, (const __half*)w11_packed, b12, b11, inv_s12, inv_s11, y_nchw,
               H, W, upsample_input ? 1 : 0, 0, 0, 0, W, W, 0, 0, (long long)H * W, nullptr, 0, 0, 0, nullptr, 0, 0, 0};
  return launch_tail_h2(h, (cudaStream_t)stream);
}

extern "C" int wctb_conv_tail_h2_sharded(const void* x_h8, const void* w12_packed, const float* b12, float inv_s12,
                                         const void* w11_packed, const float* b11, float inv_s11, int H, int W, int upsample_input,
                                         const wctb_tail_shard* sh, void* stream) {
  if (!x_h8 || !w12_packed || !b12 || !w11_packed || !b11 || !sh || !sh->out || H < 2 || W < 2) return WCTB_E_BADARG;
  if (upsample_input && ((H & 1) || (W & 1))) return WCTB_E_BADARG;
  if ((long long)H * W >= (1LL << 31)) return WCTB_E_UNSUPPORTED;
  if (sh->own_x0 < 0 || sh->own_w <= 0 || sh->own_x0 + sh->own_w > W || sh->halo < 0 || sh->halo > sh->own_w ||
      sh->out_x0 < 0 || sh->out_x0 + sh->own_w > sh->out_pitch)
    return WCTB_E_BADARG;
  if ((sh->peer_l && sh->peer_l_x0 + sh->halo > sh->peer_l_pitch) || (sh->peer_r && sh->peer_r_x0 + sh->halo > sh->peer_r_pitch))
    return WCTB_E_BADARG;
  TailH2Args h{(const uint4*)x_h8, (const __half*)w12_packed, (const __half*)w11_packed, b12, b11, inv_s12, inv_s11, sh->out,
               H, W, upsample_input ? 1 : 0, 0, 0, sh->own_x0, sh->own_w, sh->out_pitch, sh->out_x0, sh->halo,
               (long long)H * sh->out_pitch, sh->peer_l, sh->peer_l_pitch, sh->peer_l_x0, (long long)H * sh->peer_l_pitch,
               sh->peer_r, sh->peer_r_pitch, sh->peer_r_x0, (long long)H * sh->peer_r_pitch};
  return launch_tail_h2(h, (cudaStream_t)stream);
}
