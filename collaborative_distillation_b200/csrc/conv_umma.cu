// placeholder until the tcgen05 engine lands (next commit)
#include "common.cuh"
int wctb_conv3x3_p4_tf32_impl(const float*, const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t) {
  return WCTB_E_UNSUPPORTED;
}
extern "C" int wctb_pack_weights_tf32(const float*, float*, int, int, void*) { return WCTB_E_UNSUPPORTED; }
extern "C" int wctb_tf32_kgroup(int, int) { return 0; }
extern "C" int wctb_tf32_supported(int, int) { return 0; }
extern "C" int wctb_selftest_umma(float*, const float*, const float*, int, int, void*) { return WCTB_E_UNSUPPORTED; }
