// Native (no Python, starts in milliseconds) hardware check of the image-I/O kernels and the nvJPEG binding.
// Build: tests/native/build.sh  ->  tests/native/validate_io ;  run on a GPU box:  tests/native/validate_io [out.txt]
//
// It calls the C ABI exactly as the Python binding does and compares with host restatements of the same integer /
// IEEE arithmetic (the CPU suite pins those tables and formulas against PIL / torchvision, tests/test_image_io_host.py).
// Exit code 0 = every check passed.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "wctb.h"
#include "wctb_io.h"

static FILE* g_out = nullptr;
static int g_fail = 0;
#define LOG(...)                      \
  do {                                \
    printf(__VA_ARGS__);              \
    if (g_out) {                      \
      fprintf(g_out, __VA_ARGS__);    \
      fflush(g_out);                  \
    }                                 \
  } while (0)
#define CK(expr)                                                              \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      LOG("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      exit(2);                                                                \
    }                                                                         \
  } while (0)

static uint32_t rng_state = 12345u;
static inline uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

static void host_pass(const std::vector<uint8_t>& src, std::vector<uint8_t>& dst, int H, int W, int out, int axis,
                      const std::vector<int>& bounds, const std::vector<int>& coeffs, int ksize) {
  long long A = axis == 1 ? H : 1, B = axis == 1 ? 3 : 3LL * W;
  int N = axis == 1 ? W : H;
  dst.assign((size_t)(A * out * B), 0);
  for (long long a = 0; a < A; ++a)
    for (int xx = 0; xx < out; ++xx)
      for (long long b = 0; b < B; ++b) {
        int acc = 1 << 21;
        for (int i = 0; i < bounds[2 * xx + 1]; ++i)
          acc += coeffs[(size_t)xx * ksize + i] * (int)src[(size_t)((a * N + bounds[2 * xx] + i) * B + b)];
        int v = acc >> 22;
        dst[(size_t)((a * out + xx) * B + b)] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
}

static int check_resize(int H, int W, int oh, int ow) {
  std::vector<uint8_t> img((size_t)H * W * 3), ref, tmp, got;
  for (auto& v : img) v = (uint8_t)(rnd() & 255);
  uint8_t *d_a, *d_b, *d_c;
  CK(cudaMalloc(&d_a, img.size()));
  CK(cudaMalloc(&d_b, (size_t)H * ow * 3));
  CK(cudaMalloc(&d_c, (size_t)oh * ow * 3));
  CK(cudaMemcpy(d_a, img.data(), img.size(), cudaMemcpyHostToDevice));
  std::vector<uint8_t> cur = img;
  uint8_t* d_cur = d_a;
  int h = H, w = W;
  for (int step = 0; step < 2; ++step) {
    int axis = step == 0 ? 1 : 0, n_in = axis == 1 ? w : h, n_out = axis == 1 ? ow : oh;
    if (n_in == n_out) continue;
    int ksize = wctb_resize_ksize(n_in, n_out);
    std::vector<int> bounds(2 * n_out), coeffs((size_t)n_out * ksize);
    if (wctb_resize_coeffs_host(n_in, n_out, bounds.data(), coeffs.data()) != 0) return 1;
    int *d_bounds, *d_coeffs;
    CK(cudaMalloc(&d_bounds, bounds.size() * 4));
    CK(cudaMalloc(&d_coeffs, coeffs.size() * 4));
    CK(cudaMemcpy(d_bounds, bounds.data(), bounds.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_coeffs, coeffs.data(), coeffs.size() * 4, cudaMemcpyHostToDevice));
    uint8_t* d_dst = step == 0 ? d_b : d_c;
    int rc = wctb_resize_u8_pass(d_cur, d_dst, h, w, n_out, axis, d_bounds, d_coeffs, ksize, nullptr);
    if (rc != 0) { LOG("resize pass rc=%d\n", rc); return 1; }
    CK(cudaDeviceSynchronize());
    host_pass(cur, tmp, h, w, n_out, axis, bounds, coeffs, ksize);
    cur.swap(tmp);
    if (axis == 1) w = n_out; else h = n_out;
    d_cur = d_dst;
    CK(cudaFree(d_bounds));
    CK(cudaFree(d_coeffs));
  }
  got.resize(cur.size());
  CK(cudaMemcpy(got.data(), d_cur, got.size(), cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t i = 0; i < got.size(); ++i) bad += got[i] != cur[i];
  LOG("resize %dx%d -> %dx%d : %zu / %zu bytes differ  %s\n", H, W, oh, ow, bad, got.size(), bad ? "FAIL" : "ok");
  CK(cudaFree(d_a)); CK(cudaFree(d_b)); CK(cudaFree(d_c));
  return bad != 0;
}

static int check_convert(int H, int W) {
  size_t HW = (size_t)H * W;
  std::vector<uint8_t> img(HW * 3), q(HW * 3);
  for (auto& v : img) v = (uint8_t)(rnd() & 255);
  std::vector<float> f(HW * 3), x(HW * 3);
  uint8_t *d_u, *d_q;
  float *d_f, *d_x;
  CK(cudaMalloc(&d_u, HW * 3)); CK(cudaMalloc(&d_q, HW * 3)); CK(cudaMalloc(&d_f, HW * 12)); CK(cudaMalloc(&d_x, HW * 12));
  CK(cudaMemcpy(d_u, img.data(), HW * 3, cudaMemcpyHostToDevice));
  if (wctb_u8hwc_to_nchw(d_u, d_f, H, W, nullptr) != 0) return 1;
  CK(cudaMemcpy(f.data(), d_f, HW * 12, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t p = 0; p < HW; ++p)
    for (int c = 0; c < 3; ++c) {
      volatile float r = (float)img[3 * p + c] / 255.0f;
      bad += (r != f[c * HW + p]);
    }
  LOG("to_tensor %dx%d : %zu mismatches  %s\n", H, W, bad, bad ? "FAIL" : "ok");
  // quantise: random values in [-0.2, 1.4] plus exact multiples of 1/255 and half-way points
  for (size_t i = 0; i < x.size(); ++i) {
    uint32_t r = rnd();
    if ((r & 3) == 0) x[i] = (float)((r >> 2) % 256) / 255.0f;
    else if ((r & 3) == 1) x[i] = ((float)((r >> 2) % 256) + 0.5f) / 255.0f;
    else x[i] = -0.2f + 1.6f * (float)(r >> 2) / (float)(1 << 22);
  }
  CK(cudaMemcpy(d_x, x.data(), HW * 12, cudaMemcpyHostToDevice));
  if (wctb_nchw_to_u8hwc(d_x, d_q, H, W, nullptr) != 0) return 1;
  CK(cudaMemcpy(q.data(), d_q, HW * 3, cudaMemcpyDeviceToHost));
  size_t bad2 = 0;
  for (size_t p = 0; p < HW; ++p)
    for (int c = 0; c < 3; ++c) {
      volatile float m = x[c * HW + p] * 255.0f;
      volatile float s = m + 0.5f;
      float v = s < 0.f ? 0.f : (s > 255.f ? 255.f : s);
      bad2 += ((uint8_t)(int)v != q[3 * p + c]);
    }
  LOG("quantize  %dx%d : %zu mismatches  %s\n", H, W, bad2, bad2 ? "FAIL" : "ok");
  // ToTensor -> quantise is the identity on 8-bit images
  if (wctb_nchw_to_u8hwc(d_f, d_q, H, W, nullptr) != 0) return 1;
  CK(cudaMemcpy(q.data(), d_q, HW * 3, cudaMemcpyDeviceToHost));
  size_t bad3 = 0;
  for (size_t i = 0; i < HW * 3; ++i) bad3 += q[i] != img[i];
  LOG("to_tensor->quantize identity : %zu mismatches  %s\n", bad3, bad3 ? "FAIL" : "ok");
  CK(cudaFree(d_u)); CK(cudaFree(d_q)); CK(cudaFree(d_f)); CK(cudaFree(d_x));
  return (bad | bad2 | bad3) != 0;
}

static int check_jpeg(int H, int W) {
  wctb_io_codec* c = nullptr;
  int rc = wctb_io_create(&c);
  if (rc != 0) { LOG("io_create rc=%d status=%d  FAIL\n", rc, wctb_io_last_status()); return 1; }
  size_t n = (size_t)H * W * 3;
  std::vector<uint8_t> img(n), back(n);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {     // smooth synthetic picture (JPEG-friendly)
      img[(size_t)(y * W + x) * 3 + 0] = (uint8_t)(127.5 + 120 * sin(x * 0.021) * cos(y * 0.017));
      img[(size_t)(y * W + x) * 3 + 1] = (uint8_t)(127.5 + 120 * sin(x * 0.013 + y * 0.011));
      img[(size_t)(y * W + x) * 3 + 2] = (uint8_t)(255.0 * x / W * y / H);
    }
  uint8_t *d_img, *d_back;
  CK(cudaMalloc(&d_img, n)); CK(cudaMalloc(&d_back, n));
  CK(cudaMemcpy(d_img, img.data(), n, cudaMemcpyHostToDevice));
  int fail = 0;
  const int cfgs[3][2] = {{95, WCTB_IO_CSS_444}, {75, WCTB_IO_CSS_420}, {90, WCTB_IO_CSS_422}};
  for (auto& cfg : cfgs) {
    size_t len = 0, len2 = 0;
    rc = wctb_io_jpeg_encode(c, d_img, W, H, cfg[0], cfg[1], nullptr, &len);
    if (rc != 0) { LOG("jpeg_encode rc=%d status=%d  FAIL\n", rc, wctb_io_last_status()); fail = 1; continue; }
    std::vector<unsigned char> bits(len);
    rc = wctb_io_jpeg_retrieve(c, bits.data(), bits.size(), &len2, nullptr);
    if (rc != 0 || len2 != len) { LOG("jpeg_retrieve rc=%d len %zu/%zu  FAIL\n", rc, len2, len); fail = 1; continue; }
    int w = 0, h = 0, comps = 0, css = -9;
    rc = wctb_io_jpeg_info(c, bits.data(), len, &w, &h, &comps, &css);
    if (rc != 0 || w != W || h != H || comps != 3) { LOG("jpeg_info rc=%d %dx%d comps=%d  FAIL\n", rc, w, h, comps); fail = 1; continue; }
    CK(cudaMemset(d_back, 0, n));
    rc = wctb_io_jpeg_decode(c, bits.data(), len, d_back, W, H, nullptr);
    if (rc != 0) { LOG("jpeg_decode rc=%d status=%d  FAIL\n", rc, wctb_io_last_status()); fail = 1; continue; }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(back.data(), d_back, n, cudaMemcpyDeviceToHost));
    double se = 0;
    int mx = 0;
    for (size_t i = 0; i < n; ++i) {
      int d = (int)back[i] - (int)img[i];
      se += (double)d * d;
      if (abs(d) > mx) mx = abs(d);
    }
    double psnr = 10.0 * log10(255.0 * 255.0 / (se / n + 1e-12));
    bool ok = psnr > 32.0 && bits[0] == 0xFF && bits[1] == 0xD8;
    LOG("jpeg q=%d css=%d %dx%d : %zu bytes, subsampling reported %d, round-trip PSNR %.2f dB, max |d| %d  %s\n", cfg[0], cfg[1], W, H,
        len, css, psnr, mx, ok ? "ok" : "FAIL");
    fail |= !ok;
    if (g_out && cfg[0] == 75) {      // keep one bitstream for inspection with PIL back home
      FILE* f = fopen("gpurun_out/native_io_q75.jpg", "wb");
      if (f) { fwrite(bits.data(), 1, len, f); fclose(f); }
    }
  }
  wctb_io_destroy(c);
  CK(cudaFree(d_img)); CK(cudaFree(d_back));
  return fail;
}

// wctb_wct_matrix_topk with identity eigenvectors: M must be diagonal with alpha*sqrt(es_i)/sqrt(ec_i) on the directions
// kept on BOTH sides (+ (1-alpha)), which checks the rank / threshold logic on unsorted spectra with zeros.
static int check_topk(int C, int keep_c, int keep_s) {
  std::vector<double> ec(C), es(C), V((size_t)C * C, 0.0), mean(C, 0.0);
  for (int i = 0; i < C; ++i) {
    ec[i] = (i % 5 == 3) ? 0.0 : 0.01 + (double)((i * 7919) % 101);   // unsorted, distinct, a few exact zeros
    es[i] = (i % 7 == 2) ? 0.0 : 0.5 + (double)((i * 104729) % 89);
    V[(size_t)i * C + i] = 1.0;
  }
  const double tau = 1e-7, alpha = 0.75;
  double *d_ec, *d_es, *d_V, *d_mean, *d_work;
  float *d_m, *d_b, *d_mc;
  CK(cudaMalloc(&d_ec, C * 8)); CK(cudaMalloc(&d_es, C * 8)); CK(cudaMalloc(&d_V, (size_t)C * C * 8)); CK(cudaMalloc(&d_mean, C * 8));
  CK(cudaMalloc(&d_work, ((size_t)3 * C * C + 8) * 8)); CK(cudaMalloc(&d_m, (size_t)C * C * 4)); CK(cudaMalloc(&d_b, C * 4)); CK(cudaMalloc(&d_mc, C * 4));
  CK(cudaMemcpy(d_ec, ec.data(), C * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_es, es.data(), C * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_V, V.data(), (size_t)C * C * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_mean, mean.data(), C * 8, cudaMemcpyHostToDevice));
  int rc = wctb_wct_matrix_topk(d_ec, d_V, d_mean, d_es, d_V, d_mean, C, tau, alpha, keep_c, keep_s, d_m, d_b, d_mc, d_work, nullptr);
  if (rc != 0) { LOG("wct_matrix_topk rc=%d  FAIL\n", rc); return 1; }
  std::vector<float> m((size_t)C * C);
  CK(cudaMemcpy(m.data(), d_m, m.size() * 4, cudaMemcpyDeviceToHost));
  auto kept = [&](const std::vector<double>& e, int keep, int i) {
    double mx = 0;
    for (double v : e) mx = v > mx ? v : mx;
    if (!(e[i] > tau * mx)) return false;
    if (keep <= 0 || keep >= C) return true;
    int rank = 0;
    for (int j = 0; j < C; ++j) rank += e[j] > e[i];
    return rank < keep;
  };
  double worst = 0;
  int nk = 0;
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) {
      double want = 0;
      if (i == j) {
        bool k = kept(ec, keep_c, i) && kept(es, keep_s, i);
        nk += k;
        want = (k ? alpha * sqrt(es[i]) / sqrt(ec[i]) : 0.0) + (1.0 - alpha);
      }
      double d = fabs((double)m[(size_t)i * C + j] - want) / (1.0 + fabs(want));
      worst = d > worst ? d : worst;
    }
  bool ok = worst < 1e-6;
  LOG("wct_matrix_topk C=%d keep=(%d,%d): %d directions kept on both sides, worst rel err %.2e  %s\n", C, keep_c, keep_s, nk, worst, ok ? "ok" : "FAIL");
  CK(cudaFree(d_ec)); CK(cudaFree(d_es)); CK(cudaFree(d_V)); CK(cudaFree(d_mean)); CK(cudaFree(d_work)); CK(cudaFree(d_m)); CK(cudaFree(d_b)); CK(cudaFree(d_mc));
  return !ok;
}

int main(int argc, char** argv) {
  if (argc > 1) g_out = fopen(argv[1], "w");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  LOG("device: %s, %d SMs; wctb abi %d, wctb_io abi %d\n", prop.name, prop.multiProcessorCount, wctb_abi_version(), wctb_io_abi_version());
  g_fail |= check_convert(37, 53);
  g_fail |= check_convert(1080, 1920);
  g_fail |= check_resize(37, 53, 20, 28);
  g_fail |= check_resize(64, 48, 133, 100);
  g_fail |= check_resize(101, 67, 101, 33);
  g_fail |= check_resize(101, 67, 49, 67);
  g_fail |= check_resize(720, 1280, 288, 512);
  g_fail |= check_resize(2160, 3840, 1080, 1920);
  g_fail |= check_topk(24, 10, 10);
  g_fail |= check_topk(64, 30, 0);
  g_fail |= check_topk(128, 0, 0);
  g_fail |= check_topk(512, 128, 100);
  g_fail |= check_jpeg(360, 500);
  g_fail |= check_jpeg(1080, 1920);
  LOG("RESULT: %s\n", g_fail ? "FAIL" : "ALL OK");
  if (g_out) fclose(g_out);
  return g_fail;
}
