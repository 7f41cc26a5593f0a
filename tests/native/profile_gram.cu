// Minimal native driver for profiling wctb_centered_gram_fast under ncu (no Python start-up):
//   ncu --set full --clock-control none -k regex:gram_ -c 1 -o gpurun_out/gram tests/native/profile_gram [C H W]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "wctb.h"

__global__ void fill(float* p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u;
    p[i] = (float)(h >> 8) * (3.0f / 16777216.0f);
  }
}

int main(int argc, char** argv) {
  int C = argc > 3 ? atoi(argv[1]) : 24, H = argc > 3 ? atoi(argv[2]) : 2160, W = argc > 3 ? atoi(argv[3]) : 3840;
  long long n = (long long)C * H * W;
  float* x;
  double *mean, *G;
  if (cudaMalloc(&x, n * 4) || cudaMalloc(&mean, C * 8) || cudaMalloc(&G, (size_t)C * C * 8)) return 2;
  fill<<<148 * 8, 256>>>(x, n);
  cudaMemset(G, 0, (size_t)C * C * 8);
  double* hm = (double*)malloc(C * 8);
  for (int i = 0; i < C; ++i) hm[i] = 1.5;
  cudaMemcpy(mean, hm, C * 8, cudaMemcpyHostToDevice);
  int rc = wctb_centered_gram_fast(x, C, H, W, 0, H, 0, W, mean, G, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  double g00 = 0;
  cudaMemcpy(&g00, G, 8, cudaMemcpyDeviceToHost);
  printf("C=%d %dx%d rc=%d cuda=%d G[0][0]=%.6e\n", C, H, W, rc, (int)e, g00);
  return rc != 0 || e != cudaSuccess;
}
