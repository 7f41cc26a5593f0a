// libwctb: image I/O kernels around the hot path -- byte / integer work, HBM-bound.
//
//   u8 HWC -> fp32 NCHW (x / 255)                     transforms.ToTensor()      (reference data_loader.py:56-57)
//   fp32 NCHW -> u8 HWC (trunc(clamp(x*255+.5)))      vutils.save_image()        (reference WCT.py:128)
//   8-bit antialiased bilinear resize, one axis       transforms.Resize(size)    (reference data_loader.py:52-55)
//
// The resize is Pillow's 8-bit resampler restated (src/libImaging/Resample.c: precompute_coeffs,
// normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc): double-precision triangle-filter weights,
// normalised per output pixel, rounded to 22-bit fixed point on the HOST (wctb_resize_coeffs_host, no GPU needed);
// the device pass is pure int32 arithmetic, so the result is bit-exact with PIL by construction.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Resample.c

inline int grid_for(long long n, int threads) {
  long long want = (n + threads - 1) / threads;
  long long cap = (long long)wctb_num_sms() * 16;   // grid-stride: a multiple of the SM count, 16 x 256 threads resident per SM
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// One resampling pass over the middle axis of a byte tensor viewed as [A][N][B]:
//   horizontal pass of an HWC image: A = H, N = W,  B = 3
//   vertical pass:                   A = 1, N = H,  B = 3 * W   (a warp reads 32 consecutive bytes per tap)
__global__ void resize_u8_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long total, int N,
                                      long long B, int out, const int* __restrict__ bounds, const int* __restrict__ coeffs,
                                      int ksize) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long b = idx % B;
    long long t = idx / B;
    int xx = (int)(t % out);
    long long a = t / out;
    int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
    const uint8_t* s = src + (a * N + xmin) * B + b;
    const int* k = coeffs + (long long)xx * ksize;
    int acc = 1 << (PRECISION_BITS - 1);
    for (int i = 0; i < n; ++i) acc += __ldg(k + i) * (int)__ldg(s + (long long)i * B);
    int v = acc >> PRECISION_BITS;            // arithmetic shift, then clip8
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    dst[idx] = (uint8_t)v;
  }
}

__global__ void u8hwc_to_nchw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long HW) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    const uint8_t* s = src + 3 * p;
    // correctly rounded fp32 division, like torch's `img.to(float32).div(255)`
    dst[p] = __fdiv_rn((float)__ldg(s), 255.f);
    dst[HW + p] = __fdiv_rn((float)__ldg(s + 1), 255.f);
    dst[2 * HW + p] = __fdiv_rn((float)__ldg(s + 2), 255.f);
  }
}

__device__ __forceinline__ uint8_t quantize255(float x) {
  // grid.mul(255).add_(0.5).clamp_(0, 255).to(uint8): two separately rounded fp32 ops (no FMA contraction), truncation
  float v = __fadd_rn(__fmul_rn(x, 255.f), 0.5f);
  v = fminf(fmaxf(v, 0.f), 255.f);
  return (uint8_t)(int)v;
}

__global__ void nchw_to_u8hwc_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long HW) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < HW; p += (long long)gridDim.x * blockDim.x) {
    uint8_t* d = dst + 3 * p;
    d[0] = quantize255(__ldg(src + p));
    d[1] = quantize255(__ldg(src + HW + p));
    d[2] = quantize255(__ldg(src + 2 * HW + p));
  }
}

inline double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}

}  // namespace

extern "C" int wctb_resize_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return WCTB_E_BADARG;
  double filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  double support = 1.0 * filterscale;          // BILINEAR support = 1.0
  return (int)ceil(support) * 2 + 1;
}

extern "C" int wctb_resize_coeffs_host(int in_size, int out_size, int* bounds_host, int* coeffs_host) {
  if (!bounds_host || !coeffs_host) return WCTB_E_BADARG;
  int ksize = wctb_resize_ksize(in_size, out_size);
  if (ksize < 0) return ksize;
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = 1.0 * filterscale;
  double ss = 1.0 / filterscale;
  double* k = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double w = bilinear_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = xmax; x < ksize; ++x) k[x] = 0.0;
    bounds_host[2 * xx] = xmin;
    bounds_host[2 * xx + 1] = xmax;
    int* kk = coeffs_host + (long long)xx * ksize;
    for (int x = 0; x < ksize; ++x)
      kk[x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
  }
  delete[] k;
  return WCTB_OK;
}

extern "C" int wctb_resize_u8_pass(const uint8_t* src, uint8_t* dst, int H, int W, int out_size, int axis, const int* bounds,
                                   const int* coeffs, int ksize, void* stream) {
  if (!src || !dst || !bounds || !coeffs || H <= 0 || W <= 0 || out_size <= 0 || ksize <= 0 || (axis != 0 && axis != 1))
    return WCTB_E_BADARG;
  long long A = axis == 1 ? H : 1;
  int N = axis == 1 ? W : H;
  long long B = axis == 1 ? 3 : 3LL * W;
  long long total = A * out_size * B;
  resize_u8_pass_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, N, B, out_size, bounds, coeffs,
                                                                                ksize);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_u8hwc_to_nchw(const uint8_t* src_hwc, float* dst_nchw, int H, int W, void* stream) {
  if (!src_hwc || !dst_nchw || H <= 0 || W <= 0) return WCTB_E_BADARG;
  long long HW = (long long)H * W;
  u8hwc_to_nchw_kernel<<<grid_for(HW, 256), 256, 0, (cudaStream_t)stream>>>(src_hwc, dst_nchw, HW);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_nchw_to_u8hwc(const float* src_nchw, uint8_t* dst_hwc, int H, int W, void* stream) {
  if (!src_nchw || !dst_hwc || H <= 0 || W <= 0) return WCTB_E_BADARG;
  long long HW = (long long)H * W;
  nchw_to_u8hwc_kernel<<<grid_for(HW, 256), 256, 0, (cudaStream_t)stream>>>(src_nchw, dst_hwc, HW);
  WCTB_RETURN_LAUNCH();
}
