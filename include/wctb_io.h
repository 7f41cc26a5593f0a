/*
 * wctb_io.h -- C ABI of libwctb_io.so: JPEG decode / encode on the GPU (nvJPEG) for the images entering and
 * leaving the WCT stylization path.
 *
 * Replaces, around the hot path of MingSun-Tse/Collaborative-Distillation:
 *   PytorchWCT/data_loader.py:17-18,48-51   Image.open(path).convert('RGB')      -> wctb_io_jpeg_decode
 *   PytorchWCT/WCT.py:128                    vutils.save_image(img, "*.jpg")      -> wctb_io_jpeg_encode + _retrieve
 * (the pixel arithmetic between them -- resize, /255, *255+0.5 -- is in libwctb.so, see wctb.h "image I/O").
 *
 * Unlike libwctb.so this library is stateful (an nvJPEG handle with its decoder / encoder states and scratch
 * buffers lives in a `wctb_io_codec`); one codec per host thread.  JPEG entropy decoding runs on the host inside
 * nvJPEG (hybrid backend), IDCT / upsampling / colour conversion run on the GPU; the decoded image never exists
 * in host memory.  nvJPEG is library code (like cuBLAS): this file only binds it behind the repo's C ABI.
 *
 * Pixel parity with the reference's libjpeg path (PIL) is NOT bit-exact -- the JPEG standard allows +-1 per
 * IDCT and nvJPEG upsamples chroma differently from libjpeg-turbo's "fancy" filter; tests state the tolerance.
 *
 * Conventions: return 0 or a negative WCTB_IO_E_* code; never throws.  `stream` is a cudaStream_t as void*.
 */
#ifndef WCTB_IO_H_
#define WCTB_IO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WCTB_IO_ABI_VERSION 1

enum {
  WCTB_IO_OK = 0,
  WCTB_IO_E_BADARG = -1,
  WCTB_IO_E_UNSUPPORTED = -2, /* CMYK / 4-component or otherwise not decodable to RGB: caller falls back to PIL */
  WCTB_IO_E_CAPACITY = -3,    /* output buffer too small (retrieve) */
  WCTB_IO_E_CUDA = -4,
  WCTB_IO_E_NVJPEG = -5       /* see wctb_io_last_status() for the nvjpegStatus_t */
};

enum { WCTB_IO_CSS_444 = 0, WCTB_IO_CSS_422 = 1, WCTB_IO_CSS_420 = 2 };

typedef struct wctb_io_codec wctb_io_codec;

int wctb_io_abi_version(void);
const char* wctb_io_error_string(int code);
int wctb_io_last_status(void); /* nvjpegStatus_t / cudaError_t of the last failure on this thread */

int wctb_io_create(wctb_io_codec** out);
/* same with an explicit nvJPEG decode backend and flags (not yet run on hardware; wctb_io_create is the validated
 * configuration).  GPU_HYBRID moves the Huffman stage of large baseline images to the GPU (the default backend spent
 * 21 ms of host time on a 3840x2160 image); INTERP_UPSAMPLING asks nvJPEG for interpolated chroma upsampling, which is
 * closer to libjpeg's "fancy" filter that PIL uses.                                                                  */
enum { WCTB_IO_BACKEND_DEFAULT = 0, WCTB_IO_BACKEND_HYBRID = 1, WCTB_IO_BACKEND_GPU_HYBRID = 2 };
enum { WCTB_IO_FLAG_INTERP_UPSAMPLING = 1 };
int wctb_io_create_ex(int backend, unsigned flags, wctb_io_codec** out);
void wctb_io_destroy(wctb_io_codec* c);

/* header parse on the host: size, number of components (1 = grayscale, 3) and chroma subsampling (nvJPEG enum value) */
int wctb_io_jpeg_info(wctb_io_codec* c, const unsigned char* jpeg_host, size_t length, int* width, int* height,
                      int* components, int* subsampling);

/* decode to interleaved 8-bit RGB on the device: dst_hwc is [height][width][3] (pitch 3*width); grayscale files are
 * expanded to RGB like PIL's convert('RGB').  Work is enqueued on `stream` (no trailing synchronisation).          */
int wctb_io_jpeg_decode(wctb_io_codec* c, const unsigned char* jpeg_host, size_t length, uint8_t* dst_hwc, int width,
                        int height, void* stream);

/* encode a device image [height][width][3] (interleaved RGB): baseline JPEG, `quality` 1..100 (PIL's default for
 * save_image is 75), `subsampling` WCTB_IO_CSS_* (PIL's default is 4:2:0).  Synchronises `stream` and returns the
 * bitstream length; wctb_io_jpeg_retrieve then copies it to `out_host` (capacity in bytes).                          */
int wctb_io_jpeg_encode(wctb_io_codec* c, const uint8_t* src_hwc, int width, int height, int quality, int subsampling,
                        void* stream, size_t* length_out);
int wctb_io_jpeg_retrieve(wctb_io_codec* c, unsigned char* out_host, size_t capacity, size_t* length_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WCTB_IO_H_ */
