// libwctb: the whiten-and-colour transform of util_wct.py on the GPU.
//   channel sums / centred Gram (fp64 accumulate)  -> util_wct.py:68-70, 94-96
//   one-sided Jacobi eigensolver (fp64)             -> util_wct.py:74, 100 (torch.svd of a symmetric PSD matrix)
//   whitening/colouring matrix + alpha blend        -> util_wct.py:117-126, 219
//   apply  csF = M (cF - mean_c) + b                -> util_wct.py:120, 125, 126
#include <cooperative_groups.h>

#include "common.cuh"
#include "gram_small.cuh"
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------
// channel sums over a rectangular region of a P4 map
// grid (row blocks, C/4); each CTA reduces its rows of one channel chunk.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) channel_sum_kernel(const float4* __restrict__ x, int H, int W, int y0, int y1,
                                                          int x0, int x1, int rows_per_cta, double* __restrict__ out) {
  const int c4 = blockIdx.y;
  const int rbeg = y0 + blockIdx.x * rows_per_cta;
  const int rend = min(y1, rbeg + rows_per_cta);
  const float4* plane = x + (long long)c4 * H * W;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const int wreg = x1 - x0;
  for (int r = rbeg; r < rend; ++r) {
    const float4* row = plane + (long long)r * W + x0;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int n = 0;
    for (int c = threadIdx.x; c < wreg; c += 256) {
      float4 v = __ldg(row + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      if (++n == 64) { s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w; a = make_float4(0.f, 0.f, 0.f, 0.f); n = 0; }
    }
    s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
  }
  __shared__ double red[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; red[3][warp] = s3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(out + c4 * 4 + threadIdx.x, t);
  }
}
extern "C" int wctb_channel_sum(const float* x, int C, int H, int W, int y0, int y1, int x0, int x1, double* sum_out,
                                void* stream) {
  if (!x || !sum_out || C <= 0 || (C & 3) || H <= 0 || W <= 0 || y0 < 0 || y1 > H || x0 < 0 || x1 > W || y0 >= y1 ||
      x0 >= x1)
    return WCTB_E_BADARG;
  int rows = y1 - y0;
  int target = max(1, (4 * wctb_num_sms()) / (C / 4));
  int rows_per_cta = max(1, (rows + target - 1) / target);
  dim3 grid((rows + rows_per_cta - 1) / rows_per_cta, C / 4);
  channel_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)x, H, W, y0, y1, x0, x1, rows_per_cta, sum_out);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// centred Gram matrix  G += sum_p (x_p - mu)(x_p - mu)^T   in fp64 (DFMA).
// A CTA owns a GBxGB block (bi <= bj) of G and a slice of the region's pixels.  Pixels are staged centred, as
// doubles, in shared memory [pixel][GB]; SIDE x SIDE threads form a group that accumulates the whole block with a
// TR x TR register tile each (GB = SIDE*TR); 256/(SIDE*SIDE) groups take interleaved pixels of the stage, so a
// thread does GP/NG * TR*TR DFMAs between barriers.  Configurations: C=16 -> (TR 2, SIDE 8), C=24 -> (3, 8),
// C=32 -> (4, 8), C>=64 -> 64x64 blocks (4, 16) over the upper block triangle.
// Flush with fp64 atomics (order-dependent only at the 1e-16 level).
// ------------------------------------------------------------------------------------------
template <int TR, int SIDE, typename T>
__global__ void __launch_bounds__(256) centered_gram_kernel(const float4* __restrict__ x, int C, int H, int W, int y0,
                                                            int x0, int wreg, long long npix, long long pix_per_cta,
                                                            const double* __restrict__ mean, double* __restrict__ G) {
  constexpr int GB = SIDE * TR;
  constexpr int NCH4 = GB / 4;
  constexpr int TPG = SIDE * SIDE;
  constexpr int NG = 256 / TPG;
  constexpr int GP = (SIDE == 16) ? 64 : 128;      // pixels per stage
  constexpr int PITCH = GB + 2;                    // doubles; keeps 16-byte alignment of row starts
  // T = double: exact-product fp64 accumulation.  T = float (TF32 conv mode only): fp32 FFMA over the <= 32 pixels a
  // thread sees per stage, flushed into fp64 after every stage (Gram error ~1e-7 relative, far below the TF32 conv noise).
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  T(*sa)[PITCH] = reinterpret_cast<T(*)[PITCH]>(gsm_raw);
  T(*sb)[PITCH] = reinterpret_cast<T(*)[PITCH]>(gsm_raw + (size_t)GP * PITCH * sizeof(T));
  const int nb = (C + GB - 1) / GB;
  int bi = 0, bj = 0;
  {
    int t = blockIdx.y;
    for (bi = 0; bi < nb; ++bi) {
      int cnt = nb - bi;
      if (t < cnt) { bj = bi + t; break; }
      t -= cnt;
    }
  }
  const bool diag = (bi == bj);
  const long long pbeg = blockIdx.x * pix_per_cta;
  const long long pend = min(npix, pbeg + pix_per_cta);
  const int tid = threadIdx.x;
  const int grp = tid / TPG, tl = tid % TPG;
  const int ti = tl / SIDE, tj = tl % SIDE;
  T acc[TR][TR];
  double dacc[TR][TR];
#pragma unroll
  for (int u = 0; u < TR; ++u)
#pragma unroll
    for (int v = 0; v < TR; ++v) { acc[u][v] = T(0); dacc[u][v] = 0.0; }
  const long long HW = (long long)H * W;
  const int chA = bi * GB, chB = bj * GB;
  for (long long p0 = pbeg; p0 < pend; p0 += GP) {
    __syncthreads();
    for (int e = tid; e < GP * NCH4; e += 256) {
      const int pp = e % GP, ch4 = e / GP;
      const long long p = p0 + pp;
      double va[4] = {0, 0, 0, 0}, vb[4] = {0, 0, 0, 0};
      if (p < pend) {
        const int r = (int)(p / wreg), c = (int)(p - (long long)r * wreg);
        const long long off = (long long)(y0 + r) * W + (x0 + c);
        if (chA + ch4 * 4 < C) {
          const float4 v = __ldg(x + (long long)(chA / 4 + ch4) * HW + off);
          const double* m = mean + chA + ch4 * 4;
          va[0] = (double)v.x - m[0]; va[1] = (double)v.y - m[1]; va[2] = (double)v.z - m[2]; va[3] = (double)v.w - m[3];
        }
        if (!diag && chB + ch4 * 4 < C) {
          const float4 v = __ldg(x + (long long)(chB / 4 + ch4) * HW + off);
          const double* m = mean + chB + ch4 * 4;
          vb[0] = (double)v.x - m[0]; vb[1] = (double)v.y - m[1]; vb[2] = (double)v.z - m[2]; vb[3] = (double)v.w - m[3];
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) sa[pp][ch4 * 4 + k] = (T)va[k];
      if (!diag) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sb[pp][ch4 * 4 + k] = (T)vb[k];
      }
    }
    __syncthreads();
    const T(*B)[PITCH] = diag ? sa : sb;
#pragma unroll 4
    for (int pp = grp; pp < GP; pp += NG) {
      T a[TR], b[TR];
#pragma unroll
      for (int k = 0; k < TR; ++k) { a[k] = sa[pp][ti * TR + k]; b[k] = B[pp][tj * TR + k]; }
#pragma unroll
      for (int u = 0; u < TR; ++u)
#pragma unroll
        for (int v = 0; v < TR; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    if (sizeof(T) == 4) {
#pragma unroll
      for (int u = 0; u < TR; ++u)
#pragma unroll
        for (int v = 0; v < TR; ++v) { dacc[u][v] += (double)acc[u][v]; acc[u][v] = T(0); }
    }
  }
#pragma unroll
  for (int u = 0; u < TR; ++u)
#pragma unroll
    for (int v = 0; v < TR; ++v) {
      const int i = chA + ti * TR + u, j = chB + tj * TR + v;
      if (i < C && j < C) {
        const double val = dacc[u][v] + (double)acc[u][v];
        atomicAdd(G + (long long)i * C + j, val);
        if (!diag) atomicAdd(G + (long long)j * C + i, val);
      }
    }
}

template <int TR, int SIDE, typename T>
static int launch_gram(const float* x, int C, int H, int W, int y0, int y1, int x0, int x1, const double* mean,
                       double* gram_out, cudaStream_t st) {
  constexpr int GB = SIDE * TR;
  constexpr int GP = (SIDE == 16) ? 64 : 128;
  const size_t smem = (size_t)2 * GP * (GB + 2) * sizeof(T);
  static bool attr_done[WCTB_MAX_DEVICES] = {};      // per device: one process may drive several GPUs
  const int dev_slot = wctb_device_slot();
  if (!attr_done[dev_slot]) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(centered_gram_kernel<TR, SIDE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev_slot] = true;
  }
  const int nb = (C + GB - 1) / GB, nblk = nb * (nb + 1) / 2;
  const long long npix = (long long)(y1 - y0) * (x1 - x0);
  const int target = max(1, (3 * wctb_num_sms()) / nblk);
  long long per = (npix + target - 1) / target;
  per = ((per + GP - 1) / GP) * GP;
  if (per < 4 * GP) per = 4 * GP;
  dim3 grid((unsigned)((npix + per - 1) / per), nblk);
  centered_gram_kernel<TR, SIDE, T><<<grid, 256, smem, st>>>((const float4*)x, C, H, W, y0, x0, x1 - x0, npix, per, mean, gram_out);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_centered_gram(const float* x, int C, int H, int W, int y0, int y1, int x0, int x1,
                                  const double* mean, double* gram_out, void* stream) {
  if (!x || !mean || !gram_out || C <= 0 || (C & 3) || H <= 0 || W <= 0 || y0 < 0 || y1 > H || x0 < 0 || x1 > W ||
      y0 >= y1 || x0 >= x1)
    return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 16) return launch_gram<2, 8, double>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C <= 24) return launch_gram<3, 8, double>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C <= 32) return launch_gram<4, 8, double>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  return launch_gram<4, 16, double>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
}

// (the register-resident kernels for C = 24 / 32 live in gram_small.cuh, gram_ring.cu and gram_alt.cu)

// 0: register accumulation fed from a cp.async ring for C = 24 / 32 (default); 2: register accumulation fed by direct
// global loads through L1 (the first version); 1: staged shared-memory kernel everywhere; 3: variant 0 with the last
// iteration peeled out of the loop; 4: two pixels per thread and iteration + peeled (3 and 4 were written after the round's
// last GPU slot: not yet run on hardware).  tools/gram_ab.py.
static int g_gram_variant = 0;
extern "C" int wctb_debug_set_gram_variant(int v) {
  if (v < 0 || v > 4) return WCTB_E_BADARG;
  g_gram_variant = v;
  return WCTB_OK;
}

// fp32-product variant for the TF32 conv mode (see the kernel comment); same contract as wctb_centered_gram
extern "C" int wctb_centered_gram_fast(const float* x, int C, int H, int W, int y0, int y1, int x0, int x1,
                                       const double* mean, double* gram_out, void* stream) {
  if (!x || !mean || !gram_out || C <= 0 || (C & 3) || H <= 0 || W <= 0 || y0 < 0 || y1 > H || x0 < 0 || x1 > W ||
      y0 >= y1 || x0 >= x1)
    return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_gram_variant != 1 && (long long)(y1 - y0) * (x1 - x0) < (1LL << 31) - (1LL << 24)) {
    if (C == 24 || C == 32) {
      if (g_gram_variant == 0) return wctb_gram_ring_launch(C, 0, x, H, W, y0, y1, x0, x1, mean, gram_out, st);
      if (g_gram_variant == 3) return wctb_gram_ring_launch(C, 1, x, H, W, y0, y1, x0, x1, mean, gram_out, st);
      if (g_gram_variant == 4) return wctb_gram_ring2_launch(C, x, H, W, y0, y1, x0, x1, mean, gram_out, st);
      return wctb_gram_regs_launch(C, x, H, W, y0, y1, x0, x1, mean, gram_out, st);
    }
  }
  if (C <= 16) return launch_gram<2, 8, float>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C <= 24) return launch_gram<3, 8, float>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  if (C <= 32) return launch_gram<4, 8, float>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
  return launch_gram<4, 16, float>(x, C, H, W, y0, y1, x0, x1, mean, gram_out, st);
}

// ------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on G = scale*A (+I): column pairs are rotated until mutually
// orthogonal; then  sigma_k = ||g_k|| = eigenvalue,  g_k / sigma_k = eigenvector (A symmetric PSD).
// Round-robin ordering: C-1 rounds of C/2 disjoint pairs per sweep.
// Variant 1 (C <= 128): one CTA per problem, G (column-major) resident in shared memory.
// Variant 2 (any even C): cooperative grid, G in global memory (L2), grid.sync() per round.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rr_pair(int round, int k, int n, int& p, int& q) {
  // circle method: player n-1 fixed, the others rotate
  const int m = n - 1;
  if (k == 0) { p = m; q = round % m; }
  else { p = (round + k) % m; q = (round - k + m) % m; }
  if (p > q) { int t = p; p = q; q = t; }
}

struct JacobiScales { double v[8]; double early2; };   // per-problem scale; (early-stop |cos| threshold)^2 of the Cholesky-Jacobi sweeps
constexpr double JACOBI_TOL = 1e-10;   // pair converged when |g_p.g_q| <= tol |g_p||g_q|; error in eigenvalues is 2nd order
constexpr int JACOBI_MAX_SWEEPS = 40;

// rotate columns gp, gq (length n) cooperatively by LANES lanes (a power of two <= 32) of one warp
template <int LANES>
__device__ __forceinline__ int jacobi_rotate(double* gp, double* gq, int n, int sub, unsigned mask, double floor2) {
  double a = 0, b = 0, c = 0;
  for (int i = sub; i < n; i += LANES) {
    double x = gp[i], y = gq[i];
    a = fma(x, x, a); b = fma(y, y, b); c = fma(x, y, c);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(mask, a, o);
    b += __shfl_xor_sync(mask, b, o);
    c += __shfl_xor_sync(mask, c, o);
  }
  // converged pair: orthogonal to working precision, or one column is numerically null
  if (c * c <= JACOBI_TOL * JACOBI_TOL * a * b || a <= floor2 || b <= floor2) return 0;
  // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b-a)/(2c), written with one sqrt, one division, one rsqrt
  const double d = b - a, c2 = c + c;
  const double t = c2 / (d + copysign(sqrt(fma(d, d, c2 * c2)), d));
  const double cs = rsqrt(fma(t, t, 1.0));
  const double sn = cs * t;
  for (int i = sub; i < n; i += LANES) {
    double x = gp[i], y = gq[i];
    gp[i] = cs * x - sn * y;
    gq[i] = sn * x + cs * y;
  }
  return 1;
}

// Shared-memory variant (C <= 128).  One CTA per problem.
//  * channels whose diagonal entry is exactly 0 (structurally dead ReLU channels: the whole row/column of a PSD
//    matrix is then 0) are compacted away first: the solve runs on the k x k live block (k even, padded with one
//    dead channel if needed); dead channels get eigenvalue 0 and a zero eigenvector.
//  * G is column-major with a padded pitch (k+4 doubles) so concurrent column pairs spread over the banks;
//  * each column pair is owned by LANES lanes that keep both columns in registers between the dot products
//    and the rotation (EPL = elements per lane).
template <int LANES, int EPL>
__device__ __forceinline__ int jacobi_rotate_reg(double* gp, double* gq, int n, int sub, unsigned mask, double floor2) {
  double x[EPL], y[EPL];
  double a = 0, b = 0, c = 0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    int i = sub + e * LANES;
    x[e] = i < n ? gp[i] : 0.0;
    y[e] = i < n ? gq[i] : 0.0;
    a = fma(x[e], x[e], a); b = fma(y[e], y[e], b); c = fma(x[e], y[e], c);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(mask, a, o);
    b += __shfl_xor_sync(mask, b, o);
    c += __shfl_xor_sync(mask, c, o);
  }
  if (c * c <= JACOBI_TOL * JACOBI_TOL * a * b || a <= floor2 || b <= floor2) return 0;
  // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b-a)/(2c), written with one sqrt, one division, one rsqrt
  const double d = b - a, c2 = c + c;
  const double t = c2 / (d + copysign(sqrt(fma(d, d, c2 * c2)), d));
  const double cs = rsqrt(fma(t, t, 1.0));
  const double sn = cs * t;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    int i = sub + e * LANES;
    if (i < n) {
      gp[i] = cs * x[e] - sn * y[e];
      gq[i] = sn * x[e] + cs * y[e];
    }
  }
  return 1;
}

template <int LANES, int EPL>
__global__ void __launch_bounds__(64 * LANES) jacobi_smem_kernel(const double* __restrict__ A, int C,
                                                                 const JacobiScales scale, int add_identity,
                                                                 double* __restrict__ evals, double* __restrict__ evecs,
                                                                 int* __restrict__ sweeps_out) {
  extern __shared__ double G[];  // column-major k x k, pitch k+4
  __shared__ double s_fro;
  __shared__ int s_live[128];    // compacted index -> original channel
  __shared__ int s_k;
  const int prob = blockIdx.x;
  const double sc = scale.v[prob];
  const double* Ap = A + (long long)prob * C * C;
  if (threadIdx.x == 0) {
    int k = 0;
    int first_dead = -1;
    for (int i = 0; i < C; ++i) {
      double d = Ap[(long long)i * C + i] * sc + (add_identity ? 1.0 : 0.0);
      if (d > 0.0) s_live[k++] = i; else if (first_dead < 0) first_dead = i;
    }
    if ((k & 1) && first_dead >= 0) s_live[k++] = first_dead;   // keep k even for the round-robin pairing
    if (k < 2) { k = 0; }
    s_k = k;
    s_fro = 0;
  }
  __syncthreads();
  const int k = s_k;
  const int pitch = k + 4;
  for (int i = threadIdx.x; i < k * k; i += blockDim.x) {
    int r = i / k, c = i - r * k;
    int ro = s_live[r], co = s_live[c];
    double v = Ap[(long long)ro * C + co] * sc;
    if (add_identity && ro == co) v += 1.0;
    G[c * pitch + r] = v;
  }
  // outputs default: eigenvalue 0 / zero vector (dead channels, and the k == 0 case)
  for (int i = threadIdx.x; i < C; i += blockDim.x) evals[(long long)prob * C + i] = 0.0;
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) evecs[(long long)prob * C * C + i] = 0.0;
  __syncthreads();
  {
    double f = 0;
    for (int i = threadIdx.x; i < k * k; i += blockDim.x) { int r = i / k, c = i - r * k; double v = G[c * pitch + r]; f = fma(v, v, f); }
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_fro, f);
  }
  __syncthreads();
  const double floor2 = s_fro * 1e-30;  // column norm^2 below (1e-15 ||A||_F)^2 -> numerically null
  const int group = threadIdx.x / LANES, sub = threadIdx.x % LANES;
  const int ngroups = blockDim.x / LANES;
  const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));
  int sweep = 0;
  if (k >= 2) {
    for (; sweep < JACOBI_MAX_SWEEPS; ++sweep) {
      int rotated = 0;
      for (int round = 0; round < k - 1; ++round) {
        for (int pi = group; pi < k / 2; pi += ngroups) {
          int p, q;
          rr_pair(round, pi, k, p, q);
          rotated |= jacobi_rotate_reg<LANES, EPL>(G + p * pitch, G + q * pitch, k, sub, mask, floor2);
        }
        __syncthreads();
      }
      if (!__syncthreads_or(rotated)) { ++sweep; break; }
    }
  }
  // eigenvalues = column norms, eigenvectors = normalised columns, scattered back to original channel indices
  for (int j = group; j < k; j += ngroups) {
    double a = 0;
    for (int i = sub; i < k; i += LANES) a = fma(G[j * pitch + i], G[j * pitch + i], a);
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) a += __shfl_xor_sync(mask, a, o);
    double sig = sqrt(a);
    double inv = sig > 0 ? 1.0 / sig : 0.0;
    const int jo = s_live[j];
    if (sub == 0) evals[(long long)prob * C + jo] = sig;
    for (int i = sub; i < k; i += LANES)
      evecs[(long long)prob * C * C + (long long)jo * C + s_live[i]] = G[j * pitch + i] * inv;
  }
  if (sweeps_out && threadIdx.x == 0) sweeps_out[prob] = sweep;
}

// ------------------------------------------------------------------------------------------
// Variant 1b (C <= 128, the default): Jacobi on the pivoted Cholesky factor.
//   S = scale*A (+I) compacted to its live block, S = L L^T by an in-place diagonally pivoted Cholesky that stops at
//   the numerical rank (remaining diagonal <= 1e-14 max diag: the whole remaining Schur complement of a PSD matrix is
//   then below that level).  Hestenes rotations on the columns of L diagonalise L^T L, which is one Cholesky-LR step
//   closer to diagonal than S itself, so the sweep count drops from 10-12 (16-22 on synthetic spectra) to 6-9;
//   lambda_k = ||column k||^2 and the normalised columns are the eigenvectors of S.  Row order is irrelevant to the
//   column rotations, so the pivoting is virtual (a done-mask), and the L column of pivot p overwrites column p.
//   Per pair only the cross dot product is computed: the column norms^2 are tracked (a' = a - t c, b' = b + t c) and
//   recomputed exactly once per sweep; rotations are the scaled ("fast") form  x' = x - tp y, y' = y + tq x  with a
//   per-column scale s (true column = s * stored), folded back into the columns at the per-sweep refresh.
//   A sweep whose largest |cos| is below JACOBI_EARLY leaves residual cosines <= ~1e-9 (quadratic convergence) and
//   ends the iteration without a separate all-check sweep.
// ------------------------------------------------------------------------------------------
constexpr double JACOBI_EARLY = 3e-6;
constexpr double CHOL_RANK_TOL = 1e-14;

__host__ __device__ inline int jacobi_pitch(int k, int lanes) {
  // column pitch in doubles: == 8 (mod 16) for 8 lanes/pair, == 4 (mod 16) for 4 lanes/pair, so that the columns
  // touched by one half-warp request fall into disjoint banks
  return lanes >= 8 ? ((k + 7) / 16) * 16 + 8 : ((k + 11) / 16) * 16 + 4;
}

// the sweep loop of jacobi_chol_kernel for columns of at most LANES*EPL rows; returns the number of sweeps.
// Round-robin (circle method) over m = k-1 ring positions plus one fixed column.  Pair j of round r is
// {(r+j) % m, (r-j) % m}; group 0 pairs the fixed column k-1 with column r % m.  The column at ring position
// j in 1..h (h = (k-2)/2) stays in the registers of one group while it walks from position h down to 1 (its "owner");
// only the partner (the "visitor", positions -1..-h) goes through shared memory each round, which halves the
// shared-memory traffic that bounds this loop.  An owner whose column reaches position 1 stores it (it is group 0's
// visitor next round) and picks up the column entering position h, which its previous partner group has just stored.
template <int LANES, int EPL>
__device__ __forceinline__ int jacobi_chol_sweeps(double* __restrict__ G, const int k, const int pitch, const double floor2,
                                                  const double early2, double* __restrict__ s_d, double* __restrict__ s_l,
                                                  double* __restrict__ s_si) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int group = tid / LANES, sub = tid % LANES;
  const int ngroups = blockDim.x / LANES;
  const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (lane & ~(LANES - 1)));
  const int m = k - 1, h = (k - 2) / 2;
  const bool active = group < k / 2;
  int sweep = 0;
  for (; sweep < JACOBI_MAX_SWEEPS;) {
    // refresh: fold the scale into the column, exact norm^2
    for (int j = group; j < k; j += ngroups) {
      const double sj = s_l[j];
      double* g = G + j * pitch;
      double a = 0;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int i = sub + e * LANES;
        if (i < k) { const double v = g[i] * sj; g[i] = v; a = fma(v, v, a); }
      }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) a += __shfl_xor_sync(mask, a, o);
      __syncwarp(mask);   // every lane of the group has read s_l[j]
      if (sub == 0) { s_d[j] = a; s_l[j] = 1.0; s_si[j] = 1.0; }
    }
    __syncthreads();
    int big = 0;
    int own = group == 0 ? k - 1 : group;   // ring position `group` at round 0
    double x[EPL];
    if (active) {
      const double* g = G + own * pitch;
#pragma unroll
      for (int e = 0; e < EPL; ++e) x[e] = sub + e * LANES < k ? g[sub + e * LANES] : 0.0;
    }
    for (int round = 0; round < m; ++round) {
      bool handover = false;
      if (active) {
        int q = round;                                   // group 0: ring position 0
        if (group != 0) { q = 2 * round - own; if (q < 0) q += m; else if (q >= m) q -= m; }   // in (-m, 2m)
        handover = group != 0 && (own - round == 1 || own - round == 1 - m);   // ring position of own is 1
        double* gq = G + q * pitch;
        double y[EPL];
        double c0 = 0, c1 = 0;
#pragma unroll
        for (int e = 0; e < EPL; e += 2) {
          const int i0 = sub + e * LANES, i1 = i0 + LANES;
          y[e] = i0 < k ? gq[i0] : 0.0;
          y[e + 1] = i1 < k ? gq[i1] : 0.0;
          c0 = fma(x[e], y[e], c0);
          c1 = fma(x[e + 1], y[e + 1], c1);
        }
        double c = c0 + c1;
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) c += __shfl_xor_sync(mask, c, o);
        const double sp = s_l[own], sq = s_l[q];
        const double a = s_d[own], b = s_d[q];
        c *= sp * sq;
        const double cc = c * c, ab = a * b;
        const bool null = a <= floor2 || b <= floor2;
        if (!null && cc > early2 * ab) big = 1;
        if (!(null || cc <= JACOBI_TOL * JACOBI_TOL * ab)) {
          // half-angle form of the inner rotation: cos 2th = |d|/sqrt(hh), sin 2th = |2c|/sqrt(hh)
          const double d = b - a, c2 = c + c;
          const double r = rsqrt(fma(d, d, c2 * c2));
          const double cs2 = fma(0.5 * fabs(d), r, 0.5);          // cos^2 th in [1/2, 1]
          const double csi = rsqrt(cs2);                          // 1 / cos th
          const double cs = cs2 * csi;
          double t = 0.5 * fabs(c2) * r * (csi * csi);            // tan th
          if ((d < 0.0) != (c2 < 0.0)) t = -t;
          const double sip = s_si[own], siq = s_si[q];
          const double tp = t * sq * sip, tq = t * sp * siq;
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            const int i = sub + e * LANES;
            const double xn = fma(-tp, y[e], x[e]);
            if (i < k) gq[i] = fma(tq, x[e], y[e]);
            x[e] = xn;
          }
          __syncwarp(mask);   // every lane of the pair has read the scalars of own and q
          if (sub == 0) {
            s_l[own] = sp * cs; s_l[q] = sq * cs;
            s_si[own] = sip * csi; s_si[q] = siq * csi;
            s_d[own] = fma(-t, c, a); s_d[q] = fma(t, c, b);
          }
        }
        if (handover || round == m - 1) {   // the owned column becomes visible again (visitor of group 0 / end of sweep)
          double* g = G + own * pitch;
#pragma unroll
          for (int e = 0; e < EPL; ++e)
            if (sub + e * LANES < k) g[sub + e * LANES] = x[e];
        }
      }
      __syncthreads();
      if (handover && round != m - 1) {
        own = round + 1 + h; if (own >= m) own -= m;   // the column entering ring position h, stored by its last partner before the barrier
        const double* g = G + own * pitch;
#pragma unroll
        for (int e = 0; e < EPL; ++e) x[e] = sub + e * LANES < k ? g[sub + e * LANES] : 0.0;
      }
    }
    ++sweep;
    if (!__syncthreads_or(big)) break;
  }
  return sweep;
}

template <int LANES, int EPL, bool PROF>
__global__ void __launch_bounds__(64 * LANES) jacobi_chol_kernel(const double* __restrict__ A, int C,
                                                                 const JacobiScales scale, int add_identity,
                                                                 double* __restrict__ evals, double* __restrict__ evecs,
                                                                 int* __restrict__ sweeps_out, long long* __restrict__ prof) {
  // PROF: thread 0 records clock64() phase times into prof[0..3): load+compaction, Cholesky, sweeps (tools/eig_diag.py)
  long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tk = 0;
  if (PROF) tk = clock64();
#define JPROF(slot) do { if (PROF) { const long long now_ = clock64(); pt[slot] += now_ - tk; tk = now_; } } while (0)
  extern __shared__ double G[];  // column-major k x k
  __shared__ double s_d[128];    // Cholesky: running diagonal of the Schur complement; Jacobi: tracked column norms^2
  __shared__ double s_l[128];    // Cholesky: current L column; Jacobi: column scale s
  __shared__ double s_si[128];   // Jacobi: 1/s
  __shared__ int s_live[128];    // compacted index -> original channel
  __shared__ int s_done[128];
  __shared__ double s_cv[4];
  __shared__ int s_ci[4];
  __shared__ double s_tr[4];
  __shared__ int s_k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int prob = blockIdx.x;
  const double sc = scale.v[prob];
  const double* Ap = A + (long long)prob * C * C;
  const double idn = add_identity ? 1.0 : 0.0;
  // ---- live-channel compaction (warp 0, ballot prefix)
  if (warp == 0) {
    int base = 0, first_dead = -1;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int i = c0 + lane;
      const bool in = i < C;
      const double d = in ? Ap[(long long)i * C + i] * sc + idn : 0.0;
      const bool live = in && d > 0.0;
      const unsigned ml = __ballot_sync(0xffffffffu, live), md = __ballot_sync(0xffffffffu, in && !live);
      if (live) s_live[base + __popc(ml & ((1u << lane) - 1u))] = i;
      base += __popc(ml);
      if (first_dead < 0 && md) first_dead = c0 + __ffs(md) - 1;
    }
    if (lane == 0) {
      if ((base & 1) && first_dead >= 0) s_live[base++] = first_dead;   // keep k even for the round-robin pairing
      s_k = base < 2 ? 0 : base;
    }
  }
  __syncthreads();
  const int k = s_k;
  const int pitch = jacobi_pitch(k, LANES);
  for (int i = tid; i < k * k; i += blockDim.x) {
    const int r = i / k, c = i - r * k;
    const int ro = s_live[r], co = s_live[c];
    double v = Ap[(long long)ro * C + co] * sc;
    if (ro == co) { v += idn; s_d[r] = v; s_done[r] = 0; }
    G[c * pitch + r] = v;
  }
  // outputs default: eigenvalue 0 / zero vector (dead channels, null space beyond the numerical rank, k == 0)
  for (int i = tid; i < C; i += blockDim.x) evals[(long long)prob * C + i] = 0.0;
  for (int i = tid; i < C * C; i += blockDim.x) evecs[(long long)prob * C * C + i] = 0.0;
  __syncthreads();
  int sweep = 0;
  if (k >= 2) {
    // ---- trace and largest diagonal entry
    if (warp < 4) {
      double v = tid < k ? s_d[tid] : 0.0, m = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      }
      if (lane == 0) { s_tr[warp] = v; s_cv[warp] = m; }
    }
    __syncthreads();
    const double trace = (s_tr[0] + s_tr[1]) + (s_tr[2] + s_tr[3]);
    const double thr = CHOL_RANK_TOL * fmax(fmax(s_cv[0], s_cv[1]), fmax(s_cv[2], s_cv[3]));
    const double floor2 = trace * 1e-15;   // column norm^2 (= eigenvalue) below 1e-15 trace(S): numerically null
    __syncthreads();
    JPROF(0);   // load + compaction
    // ---- in-place pivoted Cholesky (two barriers per step: the pivot search of step+1 rides on the update of step)
    auto pivot_candidates = [&](double v, int idx) {   // warps 0..3: (largest remaining diagonal, smallest index) -> s_cv/s_ci
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
      }
      if (lane == 0) { s_cv[warp] = v; s_ci[warp] = idx; }
    };
    double my_d = tid < k ? s_d[tid] : -1.0;   // running Schur-complement diagonal of row tid (-1: not a candidate)
    if (warp < 4) pivot_candidates(my_d, tid);
    __syncthreads();
    for (int step = 0; step < k; ++step) {
      double dp = s_cv[0];
      int p = s_ci[0];
#pragma unroll
      for (int w = 1; w < 4; ++w) {
        const double v2 = s_cv[w];
        const int i2 = s_ci[w];
        if (v2 > dp || (v2 == dp && i2 < p)) { dp = v2; p = i2; }
      }
      if (!(dp > thr)) break;   // block-uniform: numerical rank reached
      const double inv = rsqrt(dp);
      double li = 0.0;
      if (tid < k) {
        if (tid == p) { li = dp * inv; s_done[p] = 1; my_d = -1.0; }   // s_done[p] is read by others only after the barrier
        else if (my_d >= 0.0) li = G[p * pitch + tid] * inv;           // my_d < 0 <=> row already pivoted
        G[p * pitch + tid] = li;
        s_l[tid] = li;
      }
      __syncthreads();
      {
        double lr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) lr[e] = lane + 32 * e < k ? s_l[lane + 32 * e] : 0.0;
        for (int m = warp; m < k; m += nwarps) {
          if (s_done[m]) continue;
          const double lm = s_l[m];
          double* col = G + m * pitch;
          double v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = lane + 32 * e < k ? col[lane + 32 * e] : 0.0;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (lane + 32 * e < k) col[lane + 32 * e] = fma(-lr[e], lm, v[e]);
        }
      }
      if (my_d >= 0.0) my_d = fmax(fma(-li, li, my_d), 0.0);
      if (warp < 4) pivot_candidates(my_d, tid);
      __syncthreads();
    }
    for (int m = warp; m < k; m += nwarps) {   // columns never pivoted: beyond the numerical rank
      if (s_done[m]) continue;
      for (int i = lane; i < k; i += 32) G[m * pitch + i] = 0.0;
    }
    if (tid < k) { s_l[tid] = 1.0; s_si[tid] = 1.0; }
    __syncthreads();
    JPROF(1);   // Cholesky
    // ---- Hestenes sweeps on the columns of L (instantiated for the live size: rows beyond LANES*EPL are never touched)
    if (EPL >= 16 && k <= LANES * (EPL / 4)) sweep = jacobi_chol_sweeps<LANES, (EPL >= 16 ? EPL / 4 : EPL)>(G, k, pitch, floor2, scale.early2, s_d, s_l, s_si);
    else if (k <= LANES * (EPL / 2)) sweep = jacobi_chol_sweeps<LANES, EPL / 2>(G, k, pitch, floor2, scale.early2, s_d, s_l, s_si);
    else sweep = jacobi_chol_sweeps<LANES, EPL>(G, k, pitch, floor2, scale.early2, s_d, s_l, s_si);
    const int group = tid / LANES, sub = tid % LANES;
    const int ngroups = blockDim.x / LANES;
    const unsigned mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (lane & ~(LANES - 1)));
    JPROF(2);   // sweeps
    // eigenvalues = true column norms^2, eigenvectors = normalised columns, scattered back to original channel indices
    for (int j = group; j < k; j += ngroups) {
      const double* g = G + j * pitch;
      double a = 0;
      for (int i = sub; i < k; i += LANES) a = fma(g[i], g[i], a);
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) a += __shfl_xor_sync(mask, a, o);
      const double sj = s_l[j];
      const double inv = a > 0 ? rsqrt(a) : 0.0;
      const int jo = s_live[j];
      if (sub == 0) evals[(long long)prob * C + jo] = a * sj * sj;
      for (int i = sub; i < k; i += LANES)
        evecs[(long long)prob * C * C + (long long)jo * C + s_live[i]] = g[i] * inv;
    }
  }
  if (sweeps_out && tid == 0) sweeps_out[prob] = sweep;
  if (PROF && prof && tid == 0 && prob == 0) {
    for (int i = 0; i < 8; ++i) prof[i] = pt[i];
    prof[8] = sweep;
    prof[9] = k;
  }
#undef JPROF
}

// debug: fp64 pipe probe.  out[0] = cycles of a 4096-long dependent DFMA chain (one warp); out[1] = cycles for every warp of
// a 512-thread CTA to issue 8 independent chains x 512 DFMAs (4096 per thread)
__global__ void __launch_bounds__(512) dp_rate_kernel(long long* __restrict__ out, double seed) {
  double a = seed, b = 1.0 + seed * 1e-9;
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x < 32) {
#pragma unroll 16
    for (int i = 0; i < 4096; ++i) a = fma(a, b, seed);
  }
  long long t1 = clock64();
  __syncthreads();
  double c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = seed + j;
  long long t2 = clock64();
  for (int i = 0; i < 512; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = fma(c[j], b, seed);
  }
  __syncthreads();
  long long t3 = clock64();
  double sum = a;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += c[j];
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t3 - t2; }
  if (sum == 0.123456) out[2] = 1;
}

__global__ void __launch_bounds__(256) jacobi_global_kernel(const double* __restrict__ A, int nprob, int C,
                                                            const JacobiScales scale, int add_identity,
                                                            double* __restrict__ evals, double* __restrict__ evecs,
                                                            double* __restrict__ work, int* __restrict__ flags,
                                                            int* __restrict__ sweeps_out) {
  cg::grid_group grid = cg::this_grid();
  const long long CC = (long long)C * C;
  const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long gthreads = (long long)gridDim.x * blockDim.x;
  // flags: [0..nprob) frobenius^2 (as double bits in work tail) -> use work + nprob*CC .. for fro; flags[sweep parity] rotated
  double* fro = work + nprob * CC;
  if (gtid < nprob) fro[gtid] = 0;
  if (gtid < 2) flags[gtid] = 0;
  grid.sync();
  for (long long i = gtid; i < nprob * CC; i += gthreads) {
    int prob = (int)(i / CC);
    long long e = i - prob * CC;
    int r = (int)(e / C), c = (int)(e - (long long)r * C);
    double v = A[i] * scale.v[prob];
    if (add_identity && r == c) v += 1.0;
    work[prob * CC + (long long)c * C + r] = v;
    double f = v * v;  // nprob*C*C and the stride are multiples of 32: a warp stays inside one problem
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(fro + prob, f);
  }
  grid.sync();
  const int warps_per_cta = blockDim.x >> 5;
  const int gwarp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_cta;
  const int lane = threadIdx.x & 31;
  const int pairs_per_prob = C / 2;
  int sweep = 0;
  for (; sweep < JACOBI_MAX_SWEEPS; ++sweep) {
    int rotated = 0;
    for (int round = 0; round < C - 1; ++round) {
      for (int w = gwarp; w < nprob * pairs_per_prob; w += nwarps) {
        int prob = w / pairs_per_prob, k = w - prob * pairs_per_prob;
        int p, q;
        rr_pair(round, k, C, p, q);
        double* Gp = work + prob * CC;
        rotated |= jacobi_rotate<32>(Gp + (long long)p * C, Gp + (long long)q * C, C, lane, 0xffffffffu,
                                     fro[prob] * 1e-30);
      }
      grid.sync();
    }
    if (rotated && lane == 0) atomicOr(flags + (sweep & 1), 1);
    grid.sync();
    int any = *((volatile int*)(flags + (sweep & 1)));
    if (gtid == 0) flags[(sweep + 1) & 1] = 0;
    if (!any) { ++sweep; break; }
    // the next sweep's first grid.sync orders the reset above before any atomicOr of that sweep+1 parity
  }
  grid.sync();
  for (int w = gwarp; w < nprob * C; w += nwarps) {
    int prob = w / C, k = w - prob * C;
    const double* g = work + prob * CC + (long long)k * C;
    double a = 0;
    for (int i = lane; i < C; i += 32) a = fma(g[i], g[i], a);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    double sig = sqrt(a), inv = sig > 0 ? 1.0 / sig : 0.0;
    if (lane == 0) evals[(long long)prob * C + k] = sig;
    for (int i = lane; i < C; i += 32) evecs[prob * CC + (long long)k * C + i] = g[i] * inv;
  }
  if (sweeps_out && gtid < nprob) sweeps_out[gtid] = sweep;
}

static int g_eigh_variant = 0;   // 0: Cholesky-preconditioned Jacobi (default), 1: legacy Jacobi on S (debug / A-B timing)
extern "C" int wctb_debug_set_eigh_variant(int v) {
  if (v < 0 || v > 1) return WCTB_E_BADARG;
  g_eigh_variant = v;
  return WCTB_OK;
}

static long long* g_eigh_prof = nullptr;   // debug: phase profile of the C > 64 shared-memory solve (tools/eig_diag.py)
extern "C" int wctb_debug_eigh_profile(long long* buf16) {
  g_eigh_prof = buf16;
  return WCTB_OK;
}

extern "C" int wctb_debug_dp_rate(long long* out3, void* stream) {
  if (!out3) return WCTB_E_BADARG;
  dp_rate_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(out3, 1e-3);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_eigh_jacobi(const double* a, int nprob, int C, const double* scale_host, int add_identity, double* evals,
                                double* evecs, double* work, int* sweeps_out, void* stream) {
  return wctb_eigh_jacobi_tol(a, nprob, C, scale_host, add_identity, JACOBI_EARLY, evals, evecs, work, sweeps_out, stream);
}

extern "C" int wctb_eigh_jacobi_tol(const double* a, int nprob, int C, const double* scale_host, int add_identity,
                                    double early_stop_cos, double* evals, double* evecs, double* work, int* sweeps_out,
                                    void* stream) {
  if (!a || !scale_host || !evals || !evecs || !work || nprob <= 0 || nprob > 8 || C < 2 || (C & 1)) return WCTB_E_BADARG;
  if (!(early_stop_cos >= 0.0 && early_stop_cos <= 0.1)) return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  JacobiScales scale;
  for (int i = 0; i < 8; ++i) scale.v[i] = i < nprob ? scale_host[i] : 1.0;
  scale.early2 = early_stop_cos * early_stop_cos;
  if (C <= 128 && g_eigh_variant == 0) {
    // lanes per column pair x elements per lane (LANES*EPL >= C); ~512 threads measured best on B200 (tools/eig_diag.py)
    if (C > 64) {
      const size_t smem = (size_t)C * jacobi_pitch(C, 8) * sizeof(double);
      WCTB_CUDA_TRY(cudaFuncSetAttribute(jacobi_chol_kernel<8, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (g_eigh_prof) {
        WCTB_CUDA_TRY(cudaFuncSetAttribute(jacobi_chol_kernel<8, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jacobi_chol_kernel<8, 16, true><<<nprob, 512, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out, g_eigh_prof);
        WCTB_RETURN_LAUNCH();
      }
      jacobi_chol_kernel<8, 16, false><<<nprob, 512, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out, nullptr);
    } else if (C > 32) {
      const size_t smem = (size_t)C * jacobi_pitch(C, 8) * sizeof(double);
      WCTB_CUDA_TRY(cudaFuncSetAttribute(jacobi_chol_kernel<8, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      jacobi_chol_kernel<8, 8, false><<<nprob, 512, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out, nullptr);
    } else {
      const size_t smem = (size_t)C * jacobi_pitch(C, 4) * sizeof(double);
      jacobi_chol_kernel<4, 8, false><<<nprob, 256, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out, nullptr);
    }
    WCTB_RETURN_LAUNCH();
  }
  if (C <= 128) {   // legacy variant (wctb_debug_set_eigh_variant(1)): Jacobi on the columns of S itself
    size_t smem = (size_t)C * (C + 4) * sizeof(double);
    // lanes per column pair x elements per lane (LANES*EPL >= C).  Measured on B200 (tools/eig_diag.py): ~512 threads is the
    // sweet spot -- C=128: <8,16> 1.66 ms vs <16,8> 2.12 ms; C=64: <8,8> 0.57 ms vs <4,16> 0.76 ms.
    if (C > 64) {
      WCTB_CUDA_TRY(cudaFuncSetAttribute(jacobi_smem_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      jacobi_smem_kernel<8, 16><<<nprob, 512, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out);
    } else if (C > 32) {
      WCTB_CUDA_TRY(cudaFuncSetAttribute(jacobi_smem_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      jacobi_smem_kernel<8, 8><<<nprob, 512, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out);
    } else {
      jacobi_smem_kernel<4, 8><<<nprob, 256, smem, st>>>(a, C, scale, add_identity, evals, evecs, sweeps_out);
    }
    WCTB_RETURN_LAUNCH();
  }
  // cooperative variant: work must hold nprob*C*C + nprob doubles + 2 ints (caller gives nprob*C*C + 16 doubles)
  int* flags = reinterpret_cast<int*>(work + (long long)nprob * C * C + nprob);
  int nblk_max = 0;
  WCTB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk_max, jacobi_global_kernel, 256, 0));
  if (nblk_max < 1) return WCTB_E_UNSUPPORTED;
  int want = (nprob * (C / 2) + 7) / 8;
  int gridx = min(want, wctb_num_sms() * min(nblk_max, 2));
  void* args[] = {(void*)&a, (void*)&nprob, (void*)&C, (void*)&scale, (void*)&add_identity, (void*)&evals,
                  (void*)&evecs, (void*)&work, (void*)&flags, (void*)&sweeps_out};
  WCTB_CUDA_TRY(cudaLaunchCooperativeKernel((void*)jacobi_global_kernel, dim3(gridx), dim3(256), args, 0, st));
  return WCTB_OK;
}

// ------------------------------------------------------------------------------------------
// whitening / colouring matrix (fp64), tiny: C <= 512.
//   T1 = Vc diag(Ec^-1/4 [kept]) , W = T1 T1^T ; T2 = Vs diag(Es^+1/4 [kept]), Col = T2 T2^T ; M = Col W
// evecs are stored column k at [k*C + i].
// ------------------------------------------------------------------------------------------
__global__ void eig_max_kernel(const double* __restrict__ e, int C, double* __restrict__ out) {
  double m = 0;
  for (int i = threadIdx.x; i < C; i += blockDim.x) m = fmax(m, e[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double r[32];
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmax(m, r[i]);
    *out = m;
  }
}
// out[i][j] = sum_k f(e_k) v_k[i] v_k[j]   (power = -0.5 or +0.5, thresholded at tau*emax)
// thr_abs (optional): absolute threshold written by eig_topk_threshold_kernel; a direction is then kept iff e >= *thr_abs.
__global__ void spectral_fn_kernel(const double* __restrict__ e, const double* __restrict__ v, int C, double tau,
                                   const double* __restrict__ emax, double power, double* __restrict__ out,
                                   const double* __restrict__ thr_abs = nullptr) {
  int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  __shared__ double vi[16][17], vj[16][17], f[16];
  double acc = 0;
  // default: keep e > tau*emax.  With thr_abs: keep e >= *thr_abs, written as e > (largest double below *thr_abs).
  const double thr = thr_abs ? nextafter(*thr_abs, -1.0) : tau * (*emax);
  for (int k0 = 0; k0 < C; k0 += 16) {
    int k = k0 + threadIdx.y;
    vi[threadIdx.y][threadIdx.x] = (k < C && blockIdx.y * 16 + threadIdx.x < C) ? v[(long long)k * C + blockIdx.y * 16 + threadIdx.x] : 0.0;
    vj[threadIdx.y][threadIdx.x] = (k < C && blockIdx.x * 16 + threadIdx.x < C) ? v[(long long)k * C + blockIdx.x * 16 + threadIdx.x] : 0.0;
    if (threadIdx.y == 0) {
      int kk = k0 + threadIdx.x;
      double ev = kk < C ? e[kk] : 0.0;
      f[threadIdx.x] = (kk < C && ev > thr && ev > 0) ? pow(ev, power) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc = fma(f[kk] * vi[kk][threadIdx.y], vj[kk][threadIdx.x], acc);
    __syncthreads();
  }
  if (i < C && j < C) out[(long long)i * C + j] = acc;
}
// M = alpha * (Col @ W) + (1-alpha) I  -> fp32 ; also b, mean_c
__global__ void wct_matrix_kernel(const double* __restrict__ col, const double* __restrict__ wh, int C, double alpha,
                                  const double* __restrict__ c_mean, const double* __restrict__ s_mean,
                                  float* __restrict__ m_out, float* __restrict__ b_out, float* __restrict__ mc_out,
                                  double* __restrict__ m64) {
  int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  __shared__ double a[16][17], b[16][17];
  double acc = 0;
  for (int k0 = 0; k0 < C; k0 += 16) {
    a[threadIdx.y][threadIdx.x] = (i < C && k0 + threadIdx.x < C) ? col[(long long)i * C + k0 + threadIdx.x] : 0.0;
    b[threadIdx.y][threadIdx.x] = (k0 + threadIdx.y < C && j < C) ? wh[(long long)(k0 + threadIdx.y) * C + j] : 0.0;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) acc = fma(a[threadIdx.y][kk], b[kk][threadIdx.x], acc);
    __syncthreads();
  }
  if (i < C && j < C) {
    double v = alpha * acc + (i == j ? (1.0 - alpha) : 0.0);
    m_out[(long long)i * C + j] = (float)v;
    m64[(long long)i * C + j] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && i < C) {
    b_out[i] = (float)(alpha * s_mean[i] + (1.0 - alpha) * c_mean[i]);
    mc_out[i] = (float)c_mean[i];
  }
}
// Eigenvalue truncation knobs of the reference (util_wct.py:26-27 NumEigenValue / RatEigenValue, commented uses at :87-88,
// :113-114): keep only the `keep` largest directions (and, as always, only those above tau*emax).
// thr_out = max(keep-th largest eigenvalue, smallest eigenvalue above tau*emax); +inf when nothing qualifies.  C <= 1024.
__global__ void eig_topk_threshold_kernel(const double* __restrict__ e, int C, int keep, double tau, double* __restrict__ thr_out) {
  __shared__ double se[1024];
  __shared__ double red[32];
  __shared__ double s_kth;
  const int tid = threadIdx.x, nw = blockDim.x >> 5;
  for (int i = tid; i < C; i += blockDim.x) se[i] = e[i];
  if (tid == 0) s_kth = 0.0;
  __syncthreads();
  double m = 0;
  for (int i = tid; i < C; i += blockDim.x) m = fmax(m, se[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  double emax = red[0];
  for (int i = 1; i < nw; ++i) emax = fmax(emax, red[i]);
  __syncthreads();
  const double thr_rel = tau * emax;
  const int want = (keep <= 0 || keep > C ? C : keep) - 1;     // rank (0 = largest) of the last kept eigenvalue
  double mn = INFINITY;
  for (int i = tid; i < C; i += blockDim.x) {
    const double ei = se[i];
    int rank = 0;
    for (int j = 0; j < C; ++j) rank += (se[j] > ei) || (se[j] == ei && j < i);
    if (rank == want) s_kth = ei;                              // ranks are unique: exactly one writer
    if (ei > thr_rel) mn = fmin(mn, ei);
  }
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) red[tid >> 5] = mn;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < nw; ++i) mn = fmin(mn, red[i]);
    *thr_out = fmax(s_kth, mn);
  }
}
extern "C" int wctb_wct_matrix_topk(const double* c_evals, const double* c_evecs, const double* c_mean, const double* s_evals,
                                    const double* s_evecs, const double* s_mean, int C, double tau, double alpha, int keep_c,
                                    int keep_s, float* m_out, float* b_out, float* mean_c_out, double* work, void* stream) {
  if (!c_evals || !c_evecs || !c_mean || !s_evals || !s_evecs || !s_mean || !m_out || !b_out || !mean_c_out || !work ||
      C <= 0 || C > 1024)
    return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long CC = (long long)C * C;
  double* wh = work;
  double* col = work + CC;
  double* m64 = work + 2 * CC;
  double* thr = work + 3 * CC;  // [2] absolute thresholds (content, style)
  eig_topk_threshold_kernel<<<1, 256, 0, st>>>(c_evals, C, keep_c, tau, thr);
  eig_topk_threshold_kernel<<<1, 256, 0, st>>>(s_evals, C, keep_s, tau, thr + 1);
  dim3 blk(16, 16), grd((C + 15) / 16, (C + 15) / 16);
  spectral_fn_kernel<<<grd, blk, 0, st>>>(c_evals, c_evecs, C, tau, thr, -0.5, wh, thr);
  spectral_fn_kernel<<<grd, blk, 0, st>>>(s_evals, s_evecs, C, tau, thr + 1, 0.5, col, thr + 1);
  wct_matrix_kernel<<<grd, blk, 0, st>>>(col, wh, C, alpha, c_mean, s_mean, m_out, b_out, mean_c_out, m64);
  WCTB_RETURN_LAUNCH();
}

extern "C" int wctb_wct_matrix(const double* c_evals, const double* c_evecs, const double* c_mean, const double* s_evals,
                               const double* s_evecs, const double* s_mean, int C, double tau, double alpha,
                               float* m_out, float* b_out, float* mean_c_out, double* work, void* stream) {
  if (!c_evals || !c_evecs || !c_mean || !s_evals || !s_evecs || !s_mean || !m_out || !b_out || !mean_c_out || !work ||
      C <= 0 || C > 1024)
    return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long CC = (long long)C * C;
  double* wh = work;            // W   (whitening)
  double* col = work + CC;      // Col (colouring)
  double* m64 = work + 2 * CC;  // M in fp64
  double* emax = work + 3 * CC; // [2] largest content / style eigenvalue
  eig_max_kernel<<<1, 256, 0, st>>>(c_evals, C, emax);
  eig_max_kernel<<<1, 256, 0, st>>>(s_evals, C, emax + 1);
  dim3 blk(16, 16), grd((C + 15) / 16, (C + 15) / 16);
  spectral_fn_kernel<<<grd, blk, 0, st>>>(c_evals, c_evecs, C, tau, emax, -0.5, wh);
  spectral_fn_kernel<<<grd, blk, 0, st>>>(s_evals, s_evecs, C, tau, emax + 1, 0.5, col);
  wct_matrix_kernel<<<grd, blk, 0, st>>>(col, wh, C, alpha, c_mean, s_mean, m_out, b_out, mean_c_out, m64);
  WCTB_RETURN_LAUNCH();
}

// M, b, mean_c from a ready whitening matrix (wctb_whiten_ns) and the style eigensystem
extern "C" int wctb_wct_matrix_w(const double* w_whiten, const double* c_mean, const double* s_evals, const double* s_evecs,
                                 const double* s_mean, int C, double tau, double alpha, float* m_out, float* b_out,
                                 float* mean_c_out, double* work, void* stream) {
  if (!w_whiten || !c_mean || !s_evals || !s_evecs || !s_mean || !m_out || !b_out || !mean_c_out || !work || C <= 0 || C > 1024)
    return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long CC = (long long)C * C;
  double* col = work + CC;      // same slots as wctb_wct_matrix
  double* m64 = work + 2 * CC;
  double* emax = work + 3 * CC;
  eig_max_kernel<<<1, 256, 0, st>>>(s_evals, C, emax + 1);
  dim3 blk(16, 16), grd((C + 15) / 16, (C + 15) / 16);
  spectral_fn_kernel<<<grd, blk, 0, st>>>(s_evals, s_evecs, C, tau, emax + 1, 0.5, col);
  wct_matrix_kernel<<<grd, blk, 0, st>>>(col, w_whiten, C, alpha, c_mean, s_mean, m_out, b_out, mean_c_out, m64);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// apply: y = M (x - mean_c) + b  on a P4 map.  CTA: 128 pixels x 32 output channels (blockIdx.y),
// the [C][32] slab of M^T in smem; thread = 1 pixel x 32 outputs... 2 pixels per thread to halve LDS/FMA.
// ------------------------------------------------------------------------------------------
constexpr int AP_CO = 32;
__global__ void __launch_bounds__(128) wct_apply_kernel(const float4* __restrict__ x, const float* __restrict__ m,
                                                        const float* __restrict__ b, const float* __restrict__ mean_c,
                                                        float4* __restrict__ y, int C, long long npix, int rnd) {
  extern __shared__ float4 sm4[];
  float* mt = reinterpret_cast<float*>(sm4);  // [C][AP_CO]  (input channel major)
  float* mc = mt + (size_t)C * AP_CO;         // [C]
  const int co0 = blockIdx.y * AP_CO;
  const int nco = min(AP_CO, C - co0);
  for (int i = threadIdx.x; i < C * AP_CO; i += 128) {
    int ci = i / AP_CO, o = i - ci * AP_CO;
    mt[i] = o < nco ? m[(long long)(co0 + o) * C + ci] : 0.f;
  }
  for (int i = threadIdx.x; i < C; i += 128) mc[i] = mean_c[i];
  __syncthreads();
  const long long p0 = (blockIdx.x * 128LL + threadIdx.x) * 2;
  if (p0 >= npix) return;
  const bool two = p0 + 1 < npix;
  float acc0[AP_CO], acc1[AP_CO];
#pragma unroll
  for (int o = 0; o < AP_CO; ++o) { acc0[o] = 0.f; acc1[o] = 0.f; }
  const int C4 = C >> 2;
  for (int c4 = 0; c4 < C4; ++c4) {
    float4 v0 = __ldg(x + (long long)c4 * npix + p0);
    float4 v1 = two ? __ldg(x + (long long)c4 * npix + p0 + 1) : v0;
    const float* mcc = mc + c4 * 4;
    float a0[4] = {v0.x - mcc[0], v0.y - mcc[1], v0.z - mcc[2], v0.w - mcc[3]};
    float a1[4] = {v1.x - mcc[0], v1.y - mcc[1], v1.z - mcc[2], v1.w - mcc[3]};
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float4* row = reinterpret_cast<const float4*>(mt + (size_t)(c4 * 4 + ci) * AP_CO);
#pragma unroll
      for (int o4 = 0; o4 < AP_CO / 4; ++o4) {
        float4 w = row[o4];
        acc0[o4 * 4 + 0] = fmaf(a0[ci], w.x, acc0[o4 * 4 + 0]); acc1[o4 * 4 + 0] = fmaf(a1[ci], w.x, acc1[o4 * 4 + 0]);
        acc0[o4 * 4 + 1] = fmaf(a0[ci], w.y, acc0[o4 * 4 + 1]); acc1[o4 * 4 + 1] = fmaf(a1[ci], w.y, acc1[o4 * 4 + 1]);
        acc0[o4 * 4 + 2] = fmaf(a0[ci], w.z, acc0[o4 * 4 + 2]); acc1[o4 * 4 + 2] = fmaf(a1[ci], w.z, acc1[o4 * 4 + 2]);
        acc0[o4 * 4 + 3] = fmaf(a0[ci], w.w, acc0[o4 * 4 + 3]); acc1[o4 * 4 + 3] = fmaf(a1[ci], w.w, acc1[o4 * 4 + 3]);
      }
    }
  }
#pragma unroll
  for (int o4 = 0; o4 < AP_CO / 4; ++o4) {
    if (o4 * 4 < nco) {
      const float* bb = b + co0 + o4 * 4;
      long long plane = (long long)(co0 / 4 + o4) * npix;
      float4 r0 = make_float4(acc0[o4 * 4] + bb[0], acc0[o4 * 4 + 1] + bb[1], acc0[o4 * 4 + 2] + bb[2], acc0[o4 * 4 + 3] + bb[3]);
      float4 r1 = make_float4(acc1[o4 * 4] + bb[0], acc1[o4 * 4 + 1] + bb[1], acc1[o4 * 4 + 2] + bb[2], acc1[o4 * 4 + 3] + bb[3]);
      if (rnd) {
        r0.x = wctb_tf32(r0.x); r0.y = wctb_tf32(r0.y); r0.z = wctb_tf32(r0.z); r0.w = wctb_tf32(r0.w);
        r1.x = wctb_tf32(r1.x); r1.y = wctb_tf32(r1.y); r1.z = wctb_tf32(r1.z); r1.w = wctb_tf32(r1.w);
      }
      y[plane + p0] = r0;
      if (two) y[plane + p0 + 1] = r1;
    }
  }
}
extern "C" int wctb_wct_apply(const float* x, const float* m, const float* b, const float* mean_c, float* y, int C,
                              long long npix, int round_tf32, void* stream) {
  if (!x || !m || !b || !mean_c || !y || C <= 0 || (C & 3) || C > 512 || npix <= 0) return WCTB_E_BADARG;
  size_t smem = ((size_t)C * AP_CO + C) * sizeof(float);
  WCTB_CUDA_TRY(cudaFuncSetAttribute(wct_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((npix + 255) / 256), (C + AP_CO - 1) / AP_CO);
  wct_apply_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const float4*)x, m, b, mean_c, (float4*)y, C, npix, round_tf32);
  WCTB_RETURN_LAUNCH();
}

// ------------------------------------------------------------------------------------------
// fold csF = M (x - mean_c) + b into the decoder's first conv (see wctb.h)
// ------------------------------------------------------------------------------------------
__global__ void fold_w_kernel(const float* __restrict__ w, const float* __restrict__ m, float* __restrict__ w_out,
                              int Cin, int Cout) {
  // one thread per (o, i, t); four independent fp64 accumulators (the kernel sits on the critical path of every stage and a
  // single dependent DFMA chain of Cin links is latency-bound)
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long n = (long long)Cout * Cin * 9;
  if (idx >= n) return;
  int t = (int)(idx % 9);
  int i = (int)((idx / 9) % Cin);
  int o = (int)(idx / (9LL * Cin));
  const float* wr = w + (long long)o * Cin * 9 + t;
  const float* mc = m + i;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int j = 0;
  for (; j + 3 < Cin; j += 4) {
    a0 = fma((double)wr[(long long)j * 9], (double)mc[(long long)j * Cin], a0);
    a1 = fma((double)wr[(long long)(j + 1) * 9], (double)mc[(long long)(j + 1) * Cin], a1);
    a2 = fma((double)wr[(long long)(j + 2) * 9], (double)mc[(long long)(j + 2) * Cin], a2);
    a3 = fma((double)wr[(long long)(j + 3) * 9], (double)mc[(long long)(j + 3) * Cin], a3);
  }
  for (; j < Cin; ++j) a0 = fma((double)wr[(long long)j * 9], (double)mc[(long long)j * Cin], a0);
  w_out[idx] = (float)((a0 + a1) + (a2 + a3));
}
// b_out[o] = bias[o] + sum_j (sum_t w[o][j][t]) * (b[j] - sum_i m[j][i] mean_c[i]).  One CTA: phase 1 computes the Cin
// values d[j] once (one warp per j, coalesced over i), phase 2 one warp per output channel.
__global__ void __launch_bounds__(512) fold_b_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ m,
                                                     const float* __restrict__ b, const float* __restrict__ mean_c, float* __restrict__ b_out,
                                                     int Cin, int Cout) {
  extern __shared__ double fold_d[];     // [Cin]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < Cin; j += nw) {
    double mm = 0;
    for (int i = lane; i < Cin; i += 32) mm = fma((double)m[(long long)j * Cin + i], (double)mean_c[i], mm);
    for (int s = 16; s > 0; s >>= 1) mm += __shfl_xor_sync(0xffffffffu, mm, s);
    if (lane == 0) fold_d[j] = (double)b[j] - mm;
  }
  __syncthreads();
  for (int o = warp; o < Cout; o += nw) {
    double acc = 0;
    for (int j = lane; j < Cin; j += 32) {
      const float* wp = w + ((long long)o * Cin + j) * 9;
      double ws = 0;
#pragma unroll
      for (int t = 0; t < 9; ++t) ws += (double)wp[t];
      acc = fma(ws, fold_d[j], acc);
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) b_out[o] = (float)((double)bias[o] + acc);
  }
}
extern "C" int wctb_fold_wct_into_conv(const float* w, const float* bias, const float* m, const float* b,
                                       const float* mean_c, float* w_out, float* b_out, int Cin, int Cout, void* stream) {
  if (!w || !bias || !m || !b || !mean_c || !w_out || !b_out || Cin <= 0 || Cout <= 0) return WCTB_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  long long n = (long long)Cout * Cin * 9;
  fold_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, m, w_out, Cin, Cout);
  if (Cin > 4096) return WCTB_E_UNSUPPORTED;
  fold_b_kernel<<<1, 512, (size_t)Cin * sizeof(double), st>>>(w, bias, m, b, mean_c, b_out, Cin, Cout);
  WCTB_RETURN_LAUNCH();
}
