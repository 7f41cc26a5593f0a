// libwctb: strip-halo packing for the multi-GPU path and the workspace-size query of the C ABI.
//
// A rank's image strip is NCHW [C][H][w_own]; per stage it sends its outermost `halo` columns to each neighbour and
// assembles the extended strip [C][H][lh + w_own + rh] from its own columns and what it receives.  Both directions are
// strided 2-D copies; doing them here keeps the sharded path's data movement inside the library (the Python driver
// only hands the packed buffers to NCCL).  Bit-exact by construction (plain copies).
#include "common.cuh"

namespace {

// dst[c][y][dx0 + i] = src[c][y][sx0 + i]   for i in [0, w): rows of `w` floats, one row per (c, y)
__global__ void copy_columns_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int src_pitch, int sx0,
                                    int dst_pitch, int dx0, int w) {
  const long long total = rows * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / w;
    const int i = (int)(idx - r * w);
    dst[r * dst_pitch + dx0 + i] = __ldg(src + r * src_pitch + sx0 + i);
  }
}

int launch_copy(const float* src, float* dst, long long rows, int src_pitch, int sx0, int dst_pitch, int dx0, int w, void* stream) {
  const long long total = rows * w;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)wctb_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  copy_columns_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, rows, src_pitch, sx0, dst_pitch, dx0, w);
  WCTB_RETURN_LAUNCH();
}

}  // namespace

extern "C" int wctb_halo_pack(const float* img, float* buf, int C, int H, int W, int x0, int w, void* stream) {
  if (!img || !buf || C <= 0 || H <= 0 || W <= 0 || w <= 0 || x0 < 0 || x0 + w > W) return WCTB_E_BADARG;
  return launch_copy(img, buf, (long long)C * H, W, x0, w, 0, w, stream);
}

extern "C" int wctb_halo_unpack(const float* buf, float* ext, int C, int H, int We, int x0, int w, void* stream) {
  if (!buf || !ext || C <= 0 || H <= 0 || We <= 0 || w <= 0 || x0 < 0 || x0 + w > We) return WCTB_E_BADARG;
  return launch_copy(buf, ext, (long long)C * H, w, 0, We, x0, w, stream);
}

// doubles of scratch each stateless entry point needs for `nprob` problems of size C (host; no GPU needed)
extern "C" long long wctb_workspace_doubles(int op, int C, int nprob) {
  if (C <= 0 || nprob <= 0) return WCTB_E_BADARG;
  switch (op) {
    case WCTB_WS_EIGH: return (long long)nprob * C * C + 16;        // wctb_eigh_jacobi / _tol: `work`
    case WCTB_WS_WCT_MATRIX: return 3LL * C * C + 8;                // wctb_wct_matrix / _topk / _w: `work` (nprob ignored)
    case WCTB_WS_WHITEN_NS: return 8LL * C * C + 8;                 // wctb_whiten_ns: `work` (nprob ignored)
    default: return WCTB_E_BADARG;
  }
}
