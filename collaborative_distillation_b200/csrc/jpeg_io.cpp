// libwctb_io: nvJPEG behind the C ABI of include/wctb_io.h (JPEG bytes <-> interleaved RGB u8 on the device).
// Library binding only -- the repo's own pixel kernels (resize, /255, *255+.5) are in image_io.cu / libwctb.so.
#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <new>

#include "wctb_io.h"

struct wctb_io_codec {
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t dec_state = nullptr;
  nvjpegEncoderState_t enc_state = nullptr;
  nvjpegEncoderParams_t enc_params = nullptr;
};

static thread_local int g_last_status = 0;

#define NVJ_TRY(expr)                          \
  do {                                         \
    nvjpegStatus_t s__ = (expr);               \
    if (s__ != NVJPEG_STATUS_SUCCESS) {        \
      g_last_status = (int)s__;                \
      return s__ == NVJPEG_STATUS_JPEG_NOT_SUPPORTED ? WCTB_IO_E_UNSUPPORTED : WCTB_IO_E_NVJPEG; \
    }                                          \
  } while (0)
#define CUDA_TRY(expr)                         \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) {                  \
      g_last_status = (int)e__;                \
      return WCTB_IO_E_CUDA;                   \
    }                                          \
  } while (0)

extern "C" int wctb_io_abi_version(void) { return WCTB_IO_ABI_VERSION; }
extern "C" int wctb_io_last_status(void) { return g_last_status; }
extern "C" const char* wctb_io_error_string(int code) {
  switch (code) {
    case WCTB_IO_OK: return "ok";
    case WCTB_IO_E_BADARG: return "bad argument";
    case WCTB_IO_E_UNSUPPORTED: return "JPEG variant not supported by the GPU decoder";
    case WCTB_IO_E_CAPACITY: return "output buffer too small";
    case WCTB_IO_E_CUDA: return "CUDA runtime error";
    case WCTB_IO_E_NVJPEG: return "nvJPEG error";
    default: return "unknown error";
  }
}

extern "C" void wctb_io_destroy(wctb_io_codec* c) {
  if (!c) return;
  if (c->enc_params) nvjpegEncoderParamsDestroy(c->enc_params);
  if (c->enc_state) nvjpegEncoderStateDestroy(c->enc_state);
  if (c->dec_state) nvjpegJpegStateDestroy(c->dec_state);
  if (c->handle) nvjpegDestroy(c->handle);
  delete c;
}

static int create_impl(wctb_io_codec* c, int backend, unsigned flags) {
  nvjpegBackend_t be;
  switch (backend) {
    case WCTB_IO_BACKEND_DEFAULT: be = NVJPEG_BACKEND_DEFAULT; break;
    case WCTB_IO_BACKEND_HYBRID: be = NVJPEG_BACKEND_HYBRID; break;
    case WCTB_IO_BACKEND_GPU_HYBRID: be = NVJPEG_BACKEND_GPU_HYBRID; break;
    default: return WCTB_IO_E_BADARG;
  }
  unsigned nvflags = NVJPEG_FLAGS_DEFAULT;
  if (flags & WCTB_IO_FLAG_INTERP_UPSAMPLING) nvflags |= NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION;
  if (flags & ~(unsigned)WCTB_IO_FLAG_INTERP_UPSAMPLING) return WCTB_IO_E_BADARG;
  if (backend == WCTB_IO_BACKEND_DEFAULT && nvflags == NVJPEG_FLAGS_DEFAULT) {
    NVJ_TRY(nvjpegCreateSimple(&c->handle));        // the configuration that ran on hardware in round 1
  } else {
    NVJ_TRY(nvjpegCreateEx(be, nullptr, nullptr, nvflags, &c->handle));
  }
  NVJ_TRY(nvjpegJpegStateCreate(c->handle, &c->dec_state));
  NVJ_TRY(nvjpegEncoderStateCreate(c->handle, &c->enc_state, nullptr));
  NVJ_TRY(nvjpegEncoderParamsCreate(c->handle, &c->enc_params, nullptr));
  return WCTB_IO_OK;
}

extern "C" int wctb_io_create(wctb_io_codec** out) { return wctb_io_create_ex(WCTB_IO_BACKEND_DEFAULT, 0u, out); }

extern "C" int wctb_io_create_ex(int backend, unsigned flags, wctb_io_codec** out) {
  if (!out) return WCTB_IO_E_BADARG;
  *out = nullptr;
  wctb_io_codec* c = new (std::nothrow) wctb_io_codec();
  if (!c) return WCTB_IO_E_BADARG;
  int rc = create_impl(c, backend, flags);
  if (rc != WCTB_IO_OK) {
    wctb_io_destroy(c);
    return rc;
  }
  *out = c;
  return WCTB_IO_OK;
}

extern "C" int wctb_io_jpeg_info(wctb_io_codec* c, const unsigned char* jpeg_host, size_t length, int* width, int* height,
                                 int* components, int* subsampling) {
  if (!c || !jpeg_host || length == 0 || !width || !height) return WCTB_IO_E_BADARG;
  int ncomp = 0;
  nvjpegChromaSubsampling_t css = NVJPEG_CSS_UNKNOWN;
  int ws[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0}, hs[NVJPEG_MAX_COMPONENT] = {0, 0, 0, 0};
  NVJ_TRY(nvjpegGetImageInfo(c->handle, jpeg_host, length, &ncomp, &css, ws, hs));
  *width = ws[0];
  *height = hs[0];
  if (components) *components = ncomp;
  if (subsampling) *subsampling = (int)css;
  if (ncomp != 1 && ncomp != 3) return WCTB_IO_E_UNSUPPORTED;
  return WCTB_IO_OK;
}

extern "C" int wctb_io_jpeg_decode(wctb_io_codec* c, const unsigned char* jpeg_host, size_t length, uint8_t* dst_hwc, int width,
                                   int height, void* stream) {
  if (!c || !jpeg_host || length == 0 || !dst_hwc || width <= 0 || height <= 0) return WCTB_IO_E_BADARG;
  int w = 0, h = 0, ncomp = 0;
  int rc = wctb_io_jpeg_info(c, jpeg_host, length, &w, &h, &ncomp, nullptr);
  if (rc != WCTB_IO_OK) return rc;
  if (w != width || h != height) return WCTB_IO_E_BADARG;
  nvjpegImage_t img;
  for (int i = 0; i < NVJPEG_MAX_COMPONENT; ++i) {
    img.channel[i] = nullptr;
    img.pitch[i] = 0;
  }
  img.channel[0] = dst_hwc;
  img.pitch[0] = (size_t)width * 3;
  NVJ_TRY(nvjpegDecode(c->handle, c->dec_state, jpeg_host, length, NVJPEG_OUTPUT_RGBI, &img, (cudaStream_t)stream));
  return WCTB_IO_OK;
}

extern "C" int wctb_io_jpeg_encode(wctb_io_codec* c, const uint8_t* src_hwc, int width, int height, int quality, int subsampling,
                                   void* stream, size_t* length_out) {
  if (!c || !src_hwc || width <= 0 || height <= 0 || quality < 1 || quality > 100 || !length_out) return WCTB_IO_E_BADARG;
  nvjpegChromaSubsampling_t css;
  switch (subsampling) {
    case WCTB_IO_CSS_444: css = NVJPEG_CSS_444; break;
    case WCTB_IO_CSS_422: css = NVJPEG_CSS_422; break;
    case WCTB_IO_CSS_420: css = NVJPEG_CSS_420; break;
    default: return WCTB_IO_E_BADARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  NVJ_TRY(nvjpegEncoderParamsSetQuality(c->enc_params, quality, st));
  NVJ_TRY(nvjpegEncoderParamsSetSamplingFactors(c->enc_params, css, st));
  NVJ_TRY(nvjpegEncoderParamsSetOptimizedHuffman(c->enc_params, 0, st));
  nvjpegImage_t img;
  for (int i = 0; i < NVJPEG_MAX_COMPONENT; ++i) {
    img.channel[i] = nullptr;
    img.pitch[i] = 0;
  }
  img.channel[0] = const_cast<uint8_t*>(src_hwc);
  img.pitch[0] = (size_t)width * 3;
  NVJ_TRY(nvjpegEncodeImage(c->handle, c->enc_state, c->enc_params, &img, NVJPEG_INPUT_RGBI, width, height, st));
  size_t len = 0;
  NVJ_TRY(nvjpegEncodeRetrieveBitstream(c->handle, c->enc_state, nullptr, &len, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *length_out = len;
  return WCTB_IO_OK;
}

extern "C" int wctb_io_jpeg_retrieve(wctb_io_codec* c, unsigned char* out_host, size_t capacity, size_t* length_out, void* stream) {
  if (!c || !out_host || !length_out) return WCTB_IO_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  size_t len = 0;
  NVJ_TRY(nvjpegEncodeRetrieveBitstream(c->handle, c->enc_state, nullptr, &len, st));
  if (len > capacity) {
    *length_out = len;
    return WCTB_IO_E_CAPACITY;
  }
  NVJ_TRY(nvjpegEncodeRetrieveBitstream(c->handle, c->enc_state, out_host, &len, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *length_out = len;
  return WCTB_IO_OK;
}
