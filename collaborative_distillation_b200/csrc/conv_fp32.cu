// libwctb: ABI basics, layout conversion and the CUDA-core fp32 convolution path.
//
// The fp32 path is (a) the first layer (3 input channels, K = 27: not an MMA shape) and the
// last decoder layer (3 output channels), and (b) the exact-order fp32 engine every other layer
// can be run with; the tcgen05 TF32 engine (conv_umma.cu) is validated against it.
#include "common.cuh"

thread_local int g_wctb_last_cuda_error = 0;

int wctb_conv3x3_p4_tf32_impl(const float* x, const float* w, const float* bias, float* y, int H, int W,
                              int Cin, int Cout, int epilogue, int round_tf32, cudaStream_t st);

extern "C" int wctb_abi_version(void) { return WCTB_ABI_VERSION; }
extern "C" int wctb_last_cuda_error(void) { return g_wctb_last_cuda_error; }
extern "C" const char* wctb_error_string(int code) {
  switch (code) {
    case WCTB_OK: return "ok";
    case WCTB_E_BADARG: return "bad argument";
    case WCTB_E_UNSUPPORTED: return "unsupported configuration";
    case WCTB_E_WORKSPACE: return "workspace too small";
    case WCTB_E_CUDA: return "CUDA runtime error";
    default: return "unknown error";
  }
}

// ------------------------------------------------------------------------------------------
// NCHW <-> P4
// ------------------------------------------------------------------------------------------
__global__ void nchw_to_p4_kernel(const float* __restrict__ src, float4* __restrict__ dst, int C4, long long HW, int rnd) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int c4 = blockIdx.y;
  if (p >= HW) return;
  const float* s = src + (long long)c4 * 4 * HW + p;
  float4 v = make_float4(s[0], s[HW], s[2 * HW], s[3 * HW]);
  if (rnd) { v.x = wctb_tf32(v.x); v.y = wctb_tf32(v.y); v.z = wctb_tf32(v.z); v.w = wctb_tf32(v.w); }
  dst[(long long)c4 * HW + p] = v;
}
__global__ void p4_to_nchw_kernel(const float4* __restrict__ src, float* __restrict__ dst, int C4, long long HW) {
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int c4 = blockIdx.y;
  if (p >= HW) return;
  float4 v = src[(long long)c4 * HW + p];
  float* d = dst + (long long)c4 * 4 * HW + p;
  d[0] = v.x; d[HW] = v.y; d[2 * HW] = v.z; d[3 * HW] = v.w;
}
extern "C" int wctb_nchw_to_p4(const float* src, float* dst, int C, int H, int W, int round_tf32, void* stream) {
  if (!src || !dst || C <= 0 || (C & 3) || H <= 0 || W <= 0) return WCTB_E_BADARG;
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), C / 4);
  nchw_to_p4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (float4*)dst, C / 4, HW, round_tf32);
  WCTB_RETURN_LAUNCH();
}
extern "C" int wctb_p4_to_nchw(const float* src, float* dst, int C, int H, int W, void* stream) {
  if (!src || !dst || C <= 0 || (C & 3) || H <= 0 || W <= 0) return WCTB_E_BADARG;
  long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 255) / 256), C / 4);
  p4_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, dst, C / 4, HW);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// weight packing, fp32 engine: OIHW -> [tap][Cin][Cout]
// ------------------------------------------------------------------------------------------
__global__ void pack_w_fp32_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cin, int Cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = 9 * Cin * Cout;
  if (i >= n) return;
  int co = i % Cout;
  int ci = (i / Cout) % Cin;
  int tap = i / (Cout * Cin);
  dst[i] = w[((long long)co * Cin + ci) * 9 + tap];
}
extern "C" int wctb_pack_weights_fp32(const float* w, float* dst, int Cin, int Cout, void* stream) {
  if (!w || !dst || Cin <= 0 || Cout <= 0) return WCTB_E_BADARG;
  int n = 9 * Cin * Cout;
  pack_w_fp32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, dst, Cin, Cout);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// first layer: NCHW 3 channels -> P4 Cout, conv0 folded by the host.  One pixel per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float4* __restrict__ y,
                                                         int H, int W, int Cout, int round_tf32) {
  extern __shared__ float4 smem4[];
  float* ws = reinterpret_cast<float*>(smem4);  // [27][Cout]
  float* bs = ws + 27 * Cout;
  for (int i = threadIdx.x + threadIdx.y * 32; i < 27 * Cout; i += 256) ws[i] = w[i];
  for (int i = threadIdx.x + threadIdx.y * 32; i < Cout; i += 256) bs[i] = bias[i];
  __syncthreads();
  int xx = blockIdx.x * 32 + threadIdx.x;
  int yy = blockIdx.y * 8 + threadIdx.y;
  if (xx >= W || yy >= H) return;
  float in[27];
  const long long HW = (long long)H * W;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int gy = wctb_reflect(yy + dy - 1, H);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int gx = wctb_reflect(xx + dx - 1, W);
#pragma unroll
      for (int c = 0; c < 3; ++c) in[(dy * 3 + dx) * 3 + c] = __ldg(x + c * HW + (long long)gy * W + gx);
    }
  }
  const float4* w4 = reinterpret_cast<const float4*>(ws);
  const int C4 = Cout >> 2;
  for (int c4 = 0; c4 < C4; ++c4) {
    float4 acc = reinterpret_cast<const float4*>(bs)[c4];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      float4 wv = w4[k * C4 + c4];
      acc.x = fmaf(in[k], wv.x, acc.x);
      acc.y = fmaf(in[k], wv.y, acc.y);
      acc.z = fmaf(in[k], wv.z, acc.z);
      acc.w = fmaf(in[k], wv.w, acc.w);
    }
    acc.x = wctb_relu(acc.x); acc.y = wctb_relu(acc.y); acc.z = wctb_relu(acc.z); acc.w = wctb_relu(acc.w);
    if (round_tf32) { acc.x = wctb_tf32(acc.x); acc.y = wctb_tf32(acc.y); acc.z = wctb_tf32(acc.z); acc.w = wctb_tf32(acc.w); }
    y[(long long)c4 * HW + (long long)yy * W + xx] = acc;
  }
}
// Two pixels per thread (columns xx and xx + 32 of a 64 x 8 tile): every broadcast LDS.128 of a weight quad feeds 8 FFMA
// instead of 4, which is what bounded the one-pixel kernel (27 LDS.128 per 108 FFMA); both stores stay 512-byte coalesced.
// Same arithmetic order per output as conv_first_kernel (bias, then taps 0..26), so results are bit-identical.
__global__ void __launch_bounds__(256) conv_first2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float4* __restrict__ y,
                                                          int H, int W, int Cout, int round_tf32) {
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
      const int gxb = wctb_reflect(xb + dx - 1, W);   // clamped into range when xb >= W (result unused)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ina[(dy * 3 + dx) * 3 + c] = __ldg(x + c * HW + (long long)gy * W + gxa);
        inb[(dy * 3 + dx) * 3 + c] = __ldg(x + c * HW + (long long)gy * W + gxb);
      }
    }
  }
  const float4* w4 = reinterpret_cast<const float4*>(ws);
  const int C4 = Cout >> 2;
  for (int c4 = 0; c4 < C4; ++c4) {
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
    a.x = wctb_relu(a.x); a.y = wctb_relu(a.y); a.z = wctb_relu(a.z); a.w = wctb_relu(a.w);
    b.x = wctb_relu(b.x); b.y = wctb_relu(b.y); b.z = wctb_relu(b.z); b.w = wctb_relu(b.w);
    if (round_tf32) {
      a.x = wctb_tf32(a.x); a.y = wctb_tf32(a.y); a.z = wctb_tf32(a.z); a.w = wctb_tf32(a.w);
      b.x = wctb_tf32(b.x); b.y = wctb_tf32(b.y); b.z = wctb_tf32(b.z); b.w = wctb_tf32(b.w);
    }
    float4* row = y + (long long)c4 * HW + (long long)yy * W;
    row[xa] = a;
    if (hasb) row[xb] = b;
  }
}

static int g_first_variant = 0;   // 0: two pixels per thread when W >= 64 (default), 1: one pixel per thread (A/B)
extern "C" int wctb_debug_set_first_variant(int v) {
  if (v < 0 || v > 1) return WCTB_E_BADARG;
  g_first_variant = v;
  return WCTB_OK;
}

extern "C" int wctb_conv3x3_first(const float* x, const float* w, const float* bias, float* y, int H, int W,
                                  int Cout, int round_tf32, void* stream) {
  if (!x || !w || !bias || !y || H < 2 || W < 2 || Cout <= 0 || (Cout & 3) || Cout > 512) return WCTB_E_BADARG;
  size_t smem = (size_t)(27 * Cout + Cout) * sizeof(float);
  const bool two = g_first_variant == 0 && W >= 64;
  if (smem > 48 * 1024) {
    if (two) WCTB_CUDA_TRY(cudaFuncSetAttribute(conv_first2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else WCTB_CUDA_TRY(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 block(32, 8);
  if (two) {
    dim3 grid((W + 63) / 64, (H + 7) / 8);
    conv_first2_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, w, bias, (float4*)y, H, W, Cout, round_tf32);
  } else {
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    conv_first_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(x, w, bias, (float4*)y, H, W, Cout, round_tf32);
  }
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// last decoder layer: P4 Cin -> NCHW 3 channels (+ReLU).  One pixel per thread, input tile in smem.
// ------------------------------------------------------------------------------------------
constexpr int LAST_TW = 32, LAST_TH = 8;
__global__ void __launch_bounds__(256) conv_last_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y,
                                                        int H, int W, int Cin) {
  extern __shared__ float4 smem4[];
  // weights re-laid as [tap][Cin/4][3] float4 (4 cin each); tile [LAST_TH+2][LAST_TW+2] float4 per chunk
  const int C4 = Cin >> 2;
  float4* wsm = smem4;                      // 9*C4*3
  float4* tile = smem4 + 9 * C4 * 3;        // (LAST_TH+2)*(LAST_TW+2)
  const int tid = threadIdx.x + threadIdx.y * 32;
  for (int i = tid; i < 9 * C4 * 3; i += 256) {
    int o = i % 3, c4 = (i / 3) % C4, tap = i / (3 * C4);
    const float* p = w + ((long long)tap * Cin + c4 * 4) * 3 + o;  // w[tap][ci][o]
    wsm[i] = make_float4(p[0], p[3], p[6], p[9]);
  }
  const int x0 = blockIdx.x * LAST_TW, y0 = blockIdx.y * LAST_TH;
  const int xx = x0 + threadIdx.x, yy = y0 + threadIdx.y;
  const long long HW = (long long)H * W;
  float a0 = bias[0], a1 = bias[1], a2 = bias[2];
  constexpr int TP = LAST_TW + 2;
  for (int c4 = 0; c4 < C4; ++c4) {
    __syncthreads();
    for (int i = tid; i < (LAST_TH + 2) * TP; i += 256) {
      int r = i / TP, c = i % TP;
      int gy = wctb_reflect(y0 + r - 1, H), gx = wctb_reflect(x0 + c - 1, W);
      tile[i] = __ldg(x + (long long)c4 * HW + (long long)gy * W + gx);
    }
    __syncthreads();
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float4 v = tile[(threadIdx.y + dy) * TP + threadIdx.x + dx];
        const float4* wp = wsm + ((dy * 3 + dx) * C4 + c4) * 3;
        float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        a0 = fmaf(v.x, w0.x, a0); a0 = fmaf(v.y, w0.y, a0); a0 = fmaf(v.z, w0.z, a0); a0 = fmaf(v.w, w0.w, a0);
        a1 = fmaf(v.x, w1.x, a1); a1 = fmaf(v.y, w1.y, a1); a1 = fmaf(v.z, w1.z, a1); a1 = fmaf(v.w, w1.w, a1);
        a2 = fmaf(v.x, w2.x, a2); a2 = fmaf(v.y, w2.y, a2); a2 = fmaf(v.z, w2.z, a2); a2 = fmaf(v.w, w2.w, a2);
      }
  }
  if (xx < W && yy < H) {
    long long o = (long long)yy * W + xx;
    y[o] = wctb_relu(a0);
    y[HW + o] = wctb_relu(a1);
    y[2 * HW + o] = wctb_relu(a2);
  }
}
extern "C" int wctb_conv3x3_last(const float* x, const float* w, const float* bias, float* y, int H, int W, int Cin,
                                 void* stream) {
  if (!x || !w || !bias || !y || H < 2 || W < 2 || Cin <= 0 || (Cin & 3) || Cin > 512) return WCTB_E_BADARG;
  dim3 grid((W + LAST_TW - 1) / LAST_TW, (H + LAST_TH - 1) / LAST_TH), block(32, 8);
  size_t smem = (size_t)(9 * (Cin / 4) * 3 + (LAST_TH + 2) * (LAST_TW + 2)) * sizeof(float4);
  if (smem > 48 * 1024) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(conv_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  conv_last_kernel<<<grid, block, smem, (cudaStream_t)stream>>>((const float4*)x, w, bias, y, H, W, Cin);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// middle layers, fp32 engine.  CTA = 32x32 output pixels x 8 output channels; each thread a 2x2
// pixel quad x 8 channels (32 accumulators).  Input chunk (4 channels) staged in smem per step.
// Accumulation order per output: channel-chunk major, then tap, then channel (fixed, so results
// do not depend on the tile or the GPU a pixel lands on).
// ------------------------------------------------------------------------------------------
constexpr int FT = 32;            // tile edge (pixels)
constexpr int FTP = FT + 2;       // with halo
template <int EPI>
__global__ void __launch_bounds__(256) conv_p4_fp32_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float4* __restrict__ y,
                                                           int H, int W, int Cin, int Cout, int round_tf32) {
  __shared__ float4 tile[FTP * FTP];
  __shared__ float4 wsm[9 * 4 * 2];  // [tap][ci][2 x float4 = 8 couts]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int x0 = blockIdx.x * FT, y0 = blockIdx.y * FT;
  const int g = blockIdx.z;  // cout group of 8
  const long long HW = (long long)H * W;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[p][o] = 0.f;

  const int C4 = Cin >> 2;
  for (int c4 = 0; c4 < C4; ++c4) {
    __syncthreads();
    for (int i = tid; i < FTP * FTP; i += 256) {
      int r = i / FTP, c = i - r * FTP;
      int gy = wctb_reflect(y0 + r - 1, H), gx = wctb_reflect(x0 + c - 1, W);
      tile[i] = __ldg(x + (long long)c4 * HW + (long long)gy * W + gx);
    }
    if (tid < 72) {
      int half = tid & 1, ci = (tid >> 1) & 3, tap = tid >> 3;
      wsm[tid] = __ldg(reinterpret_cast<const float4*>(w + ((long long)tap * Cin + c4 * 4 + ci) * Cout + g * 8) + half);
    }
    __syncthreads();
    float4 in[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) in[r][c] = tile[(2 * ty + r) * FTP + 2 * tx + c];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          float4 wa = wsm[((dy * 3 + dx) * 4 + ci) * 2], wb = wsm[((dy * 3 + dx) * 4 + ci) * 2 + 1];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float4 v4 = in[(p >> 1) + dy][(p & 1) + dx];
            float v = ci == 0 ? v4.x : ci == 1 ? v4.y : ci == 2 ? v4.z : v4.w;
            acc[p][0] = fmaf(v, wa.x, acc[p][0]); acc[p][1] = fmaf(v, wa.y, acc[p][1]);
            acc[p][2] = fmaf(v, wa.z, acc[p][2]); acc[p][3] = fmaf(v, wa.w, acc[p][3]);
            acc[p][4] = fmaf(v, wb.x, acc[p][4]); acc[p][5] = fmaf(v, wb.y, acc[p][5]);
            acc[p][6] = fmaf(v, wb.z, acc[p][6]); acc[p][7] = fmaf(v, wb.w, acc[p][7]);
          }
        }
  }
  // epilogue: bias + ReLU (+ TF32 rounding for a tensor-core consumer)
  float bv[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) bv[o] = __ldg(bias + g * 8 + o);
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float v = wctb_relu(acc[p][o] + bv[o]);
      acc[p][o] = round_tf32 ? wctb_tf32(v) : v;
    }
  const int py0 = y0 + 2 * ty, px0 = x0 + 2 * tx;
  if (EPI == WCTB_EPI_NONE) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int yy = py0 + (p >> 1), xx = px0 + (p & 1);
      if (yy < H && xx < W) {
        long long o = (long long)yy * W + xx;
        y[(long long)(2 * g) * HW + o] = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        y[(long long)(2 * g + 1) * HW + o] = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
      }
    }
  } else if (EPI == WCTB_EPI_POOL2) {
    const int Ho = H >> 1, Wo = W >> 1;
    int oy = py0 >> 1, ox = px0 >> 1;
    if (oy < Ho && ox < Wo) {
      float m[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) m[o] = fmaxf(fmaxf(acc[0][o], acc[1][o]), fmaxf(acc[2][o], acc[3][o]));
      long long o = (long long)oy * Wo + ox, HWo = (long long)Ho * Wo;
      y[(long long)(2 * g) * HWo + o] = make_float4(m[0], m[1], m[2], m[3]);
      y[(long long)(2 * g + 1) * HWo + o] = make_float4(m[4], m[5], m[6], m[7]);
    }
  } else {  // WCTB_EPI_UP2
    const int Wo = W * 2;
    const long long HWo = HW * 4;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int yy = py0 + (p >> 1), xx = px0 + (p & 1);
      if (yy < H && xx < W) {
        float4 lo = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        float4 hi = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            long long o = (long long)(2 * yy + a) * Wo + 2 * xx + b;
            y[(long long)(2 * g) * HWo + o] = lo;
            y[(long long)(2 * g + 1) * HWo + o] = hi;
          }
      }
    }
  }
}

extern "C" int wctb_conv3x3_p4(const float* x, const float* w, const float* bias, float* y, int H, int W, int Cin,
                               int Cout, int epilogue, int round_tf32, int engine, void* stream) {
  if (!x || !w || !bias || !y || H < 2 || W < 2 || Cin <= 0 || Cout <= 0 || (Cin & 3) || (Cout & 7))
    return WCTB_E_BADARG;
  if (epilogue < 0 || epilogue > 3 || (epilogue == WCTB_EPI_NCHW3 && engine != WCTB_ENGINE_TF32)) return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == WCTB_ENGINE_TF32) return wctb_conv3x3_p4_tf32_impl(x, w, bias, y, H, W, Cin, Cout, epilogue, round_tf32, st);
  if (engine != WCTB_ENGINE_FP32) return WCTB_E_BADARG;
  dim3 grid((W + FT - 1) / FT, (H + FT - 1) / FT, Cout / 8);
  if (grid.y > 65535 || grid.z > 65535) return WCTB_E_UNSUPPORTED;
  const float4* x4 = (const float4*)x;
  float4* y4 = (float4*)y;
  switch (epilogue) {
    case WCTB_EPI_NONE: conv_p4_fp32_kernel<WCTB_EPI_NONE><<<grid, 256, 0, st>>>(x4, w, bias, y4, H, W, Cin, Cout, round_tf32); break;
    case WCTB_EPI_POOL2: conv_p4_fp32_kernel<WCTB_EPI_POOL2><<<grid, 256, 0, st>>>(x4, w, bias, y4, H, W, Cin, Cout, round_tf32); break;
    default: conv_p4_fp32_kernel<WCTB_EPI_UP2><<<grid, 256, 0, st>>>(x4, w, bias, y4, H, W, Cin, Cout, round_tf32); break;
  }
  WCTB_RETURN_LAUNCH();
}
