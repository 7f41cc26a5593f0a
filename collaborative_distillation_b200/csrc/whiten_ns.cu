// libwctb: whitening matrix W = S^-1/2 (pseudo-inverse on the range of S) WITHOUT an eigendecomposition -- GEMMs only, so
// that the C <= 128 problem on the critical path spreads over a cooperative grid instead of one CTA.
//
//   replaces (content side): torch.svd(contentConv) + c_v diag(c_e^-1/2) c_v^T      (reference util_wct.py:74, 117-119)
//
//   S = scale * G (+ I)                     centred Gram of the content features (wctb_centered_gram*)
//   S = L L^T                               rank-revealing pivoted Cholesky, L: C x r  (one CTA, shared memory); dead
//                                           channels and numerically null directions never become pivots -> zero rows
//   B = L^T L                               r x r, full rank, cond(B) = cond(S on its range)
//   R0 = B / s, Z0 = I                      s = max row sum of |B| >= lambda_max
//   T = (3I - R)/2;  Z <- T Z;  R <- T R T  coupled Newton-Schulz: R -> I, Z -> (B/s)^-1/2; stop when max|I - R| < 1e-13
//   W = L Z^3 L^T / s^1.5                   since (L B^-3/2 L^T)^2 = L B^-2 L^T = S^+
//
// Every step after the Cholesky is an r^3 GEMM (<= 2 MFLOP) split by output rows over the CTAs, operands in L2, one
// grid.sync() per dependent GEMM.  tools/ns_invsqrt_prototype.py is the numpy model of exactly this sequence: 11-19
// iterations on the cfg3 spectra (26 at cond 1e7), agreement with LAPACK's pseudo-inverse square root 1e-14..2e-13, and
// the reference's whiten_and_color goldens reproduced to 1e-15 (3.7e-8 on the rank-deficient HW < C case).
//
// STATUS: written after the last GPU slot of round 1 -- compiles for sm_100a, not yet run on hardware.  Opt-in
// (WCT.whiten_solver = "ns" / WCTB_WHITEN=ns); its tests carry the pending_hw marker.  The Jacobi solver stays the default.
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

namespace {

constexpr int NS_THREADS = 256;
constexpr int NS_CTAS = 16;          // 8 output rows per CTA at C = 128
constexpr int NS_ROWS = 8;           // rows of a GEMM's output one CTA owns (two groups of 4 accumulators)
constexpr int NS_MAX_C = 128;
constexpr int NS_MAXIT = 40;
constexpr double NS_TOL = 1e-13;
constexpr double NS_RANK_CUT = 1e-10;   // pivot <= cut * largest diagonal entry: numerically null (prototype: any cut in
                                        // [1e-14, 1e-7] gives the same goldens)

// ---- pivoted Cholesky of S (C x C, symmetric PSD) by ONE CTA in shared memory.  Writes L (C x r, pitch C) and L^T
// (r x C, pitch C) to global memory and returns r.  Row order is never permuted (virtual pivoting with a done-mask):
// column j of L belongs to the j-th pivot; rows that were pivots earlier are zero in later columns.
__device__ int ns_cholesky(const double* gram, double scale, double idn, int C, double* S, double* l, int* done, double* red_v,
                           int* red_i, double* L, double* LT) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C * C; i += NS_THREADS) {
    const int r = i / C, c = i - r * C;
    S[i] = gram[i] * scale + (r == c ? idn : 0.0);
    L[i] = 0.0;
    LT[i] = 0.0;
  }
  if (tid < C) done[tid] = 0;
  __syncthreads();
  double thr = 0.0;
  int rank = 0;
  for (int j = 0; j < C; ++j) {
    // largest remaining diagonal entry (ties: lowest index) -- C <= 128: warps 0..3 hold one candidate per lane
    double v = -1.0;
    int idx = 0x7fffffff;
    if (tid < C && !done[tid]) { v = S[tid * C + tid]; idx = tid; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { red_v[warp] = v; red_i[warp] = idx; }
    __syncthreads();
    double best = red_v[0];
    int p = red_i[0];
#pragma unroll
    for (int w = 1; w < NS_THREADS / 32; ++w)
      if (red_v[w] > best || (red_v[w] == best && red_i[w] < p)) { best = red_v[w]; p = red_i[w]; }
    if (j == 0) thr = NS_RANK_CUT * best;
    if (!(best > thr) || !(best > 0.0)) break;       // uniform: every thread sees the same reduction result
    const double piv = sqrt(best);
    if (tid < C) {
      const double lv = done[tid] ? 0.0 : (tid == p ? piv : S[tid * C + p] / piv);
      l[tid] = lv;
      L[tid * C + j] = lv;
      LT[j * C + tid] = lv;
    }
    __syncthreads();                                 // l complete; every thread has read done[] and red_*[]
    if (tid == 0) done[p] = 1;
    // Schur complement: S -= l l^T everywhere (rows that are done have l = 0; row/column p is never read again)
    for (int i = tid; i < C * C; i += NS_THREADS) {
      const int r = i / C, c = i - r * C;
      S[i] = fma(-l[r], l[c], S[i]);
    }
    __syncthreads();
    rank = j + 1;
  }
  return rank;
}

// ---- out[M x N] = alpha * opA(A)[M x K] * opB(B)[K x N], all matrices row-major with pitch ld, rows split over the grid.
// op: 0 = plain, 1 = the Newton-Schulz factor T = 1.5 I - 0.5 X formed on the fly.  A CTA owns NS_ROWS consecutive output
// rows.  It first copies ALL of B (K x N doubles <= 128 KB, coalesced, every load in flight at once) and its A rows into
// shared memory, then thread (n, h) accumulates rows 4h..4h+3 of column n from there (A: broadcast reads, B: conflict-free
// along n).  The buffers are rewritten between grid.sync()s by other CTAs: plain (coherent) loads only, no __ldg.
template <int AOP, int BOP>
__device__ void ns_gemm(double* out, const double* A, const double* B, int M, int N, int K, int ld, double alpha, double* sA,
                        double* sB) {
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * NS_ROWS;               // gridDim.x * NS_ROWS >= 128 >= M: one tile per CTA
  if (m0 >= M) return;                               // uniform per CTA (the barriers below are CTA-wide only)
  __syncthreads();                                   // sA / sB free (previous call)
  for (int i = tid; i < K * N; i += NS_THREADS) {
    const int k = i / N, n = i - k * N;
    double b = B[k * ld + n];
    if (BOP == 1) b = (k == n ? 1.5 : 0.0) - 0.5 * b;
    sB[i] = b;
  }
  for (int i = tid; i < NS_ROWS * K; i += NS_THREADS) {
    const int rr = i / K, k = i - rr * K;
    const int m = m0 + rr;
    double a = 0.0;
    if (m < M) {
      a = A[m * ld + k];
      if (AOP == 1) a = (m == k ? 1.5 : 0.0) - 0.5 * a;
    }
    sA[k * NS_ROWS + rr] = a;                        // [k][row]: the 4 rows of a thread are 32 contiguous bytes
  }
  __syncthreads();
  const int n = tid & (NS_MAX_C - 1), h = tid >> 7;
  if (n < N) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const double2* a2 = reinterpret_cast<const double2*>(sA + h * 4);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double bv = sB[k * N + n];
      const double2 a01 = a2[k * (NS_ROWS / 2)], a23 = a2[k * (NS_ROWS / 2) + 1];   // broadcast LDS.128 x 2
      acc[0] = fma(a01.x, bv, acc[0]);
      acc[1] = fma(a01.y, bv, acc[1]);
      acc[2] = fma(a23.x, bv, acc[2]);
      acc[3] = fma(a23.y, bv, acc[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + h * 4 + i;
      if (m < M) out[m * ld + n] = alpha * acc[i];
    }
  }
}

// block-wide max of a non-negative value, result to every thread
__device__ double ns_block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = red[0];
#pragma unroll
  for (int w = 1; w < NS_THREADS / 32; ++w) m = fmax(m, red[w]);
  return m;
}

__global__ void __launch_bounds__(NS_THREADS) whiten_ns_kernel(const double* gram, double scale, int add_identity, int C,
                                                               double* w_out, double* work, int* info_out) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double ns_smem[];               // C*C doubles: S of the Cholesky (CTA 0), then the B operand of each GEMM
  __shared__ __align__(16) double sA[NS_ROWS * NS_MAX_C];   // [k][row]
  __shared__ double s_l[NS_MAX_C];
  __shared__ int s_done[NS_MAX_C];
  __shared__ double s_red[NS_THREADS / 32];
  __shared__ int s_redi[NS_THREADS / 32];
  const int CC = C * C;
  double* L = work;
  double* LT = work + CC;
  double* R = work + 2 * CC;
  double* R2 = work + 3 * CC;
  double* Z = work + 4 * CC;
  double* Z2 = work + 5 * CC;
  double* U = work + 6 * CC;
  double* P = work + 7 * CC;
  int* g_rank = reinterpret_cast<int*>(work + 8 * CC);
  const int tid = threadIdx.x;

  if (blockIdx.x == 0) {
    const int r = ns_cholesky(gram, scale, add_identity ? 1.0 : 0.0, C, ns_smem, s_l, s_done, s_red, s_redi, L, LT);
    if (tid == 0) *g_rank = r;
  }
  grid.sync();
  const int r = *reinterpret_cast<volatile int*>(g_rank);
  if (r == 0) {                                      // S == 0: W = 0
    for (int i = blockIdx.x * NS_THREADS + tid; i < CC; i += gridDim.x * NS_THREADS) w_out[i] = 0.0;
    if (info_out && blockIdx.x == 0 && tid == 0) { info_out[0] = 0; info_out[1] = 0; info_out[2] = 1; }
    return;                                          // uniform over the grid: no later grid.sync is skipped by a subset
  }

  // B = L^T L  (into R2), then s = max_i sum_j |B_ij| (every CTA computes it redundantly: identical bits everywhere)
  ns_gemm<0, 0>(R2, LT, L, r, r, C, C, 1.0, sA, ns_smem);
  grid.sync();
  double rowsum = 0.0;
  if (tid < r)
    for (int j = 0; j < r; ++j) rowsum += fabs(R2[tid * C + j]);
  const double s = ns_block_max(rowsum, s_red);
  const double inv_s = 1.0 / s;
  for (int i = blockIdx.x * NS_THREADS + tid; i < r * r; i += gridDim.x * NS_THREADS) {
    const int a = i / r, b = i - a * r;
    R[a * C + b] = R2[a * C + b] * inv_s;
    Z[a * C + b] = (a == b) ? 1.0 : 0.0;
  }
  grid.sync();

  int it = 0;
  double err = 0.0;
  for (;; ++it) {
    // residual max|I - R| from the whole matrix, redundantly per CTA (r*r <= 16K loads): a uniform decision without a
    // reduction across the grid
    double e = 0.0;
    for (int i = tid; i < r * r; i += NS_THREADS) {
      const int a = i / r, b = i - a * r;
      e = fmax(e, fabs((a == b ? 1.0 : 0.0) - R[a * C + b]));
    }
    err = ns_block_max(e, s_red);
    if (err < NS_TOL || it == NS_MAXIT) break;
    ns_gemm<1, 0>(Z2, R, Z, r, r, r, C, 1.0, sA, ns_smem);    // Z' = T Z
    ns_gemm<1, 0>(U, R, R, r, r, r, C, 1.0, sA, ns_smem);     // U  = T R
    grid.sync();
    ns_gemm<0, 1>(R2, U, R, r, r, r, C, 1.0, sA, ns_smem);    // R' = U T
    grid.sync();
    double* t = R; R = R2; R2 = t;
    t = Z; Z = Z2; Z2 = t;
  }

  // W = L Z^3 L^T / s^1.5
  ns_gemm<0, 0>(U, Z, Z, r, r, r, C, 1.0, sA, ns_smem);       // Z^2
  grid.sync();
  ns_gemm<0, 0>(R2, U, Z, r, r, r, C, 1.0, sA, ns_smem);      // Z^3
  grid.sync();
  ns_gemm<0, 0>(P, L, R2, C, r, r, C, 1.0, sA, ns_smem);      // L Z^3        (C x r)
  grid.sync();
  ns_gemm<0, 0>(w_out, P, LT, C, C, r, C, inv_s * sqrt(inv_s), sA, ns_smem);   // (L Z^3) L^T / s^1.5   (C x C)
  if (info_out && blockIdx.x == 0 && tid == 0) {
    info_out[0] = r;
    info_out[1] = it;
    info_out[2] = err < NS_TOL ? 1 : 0;
  }
}

}  // namespace

extern "C" int wctb_whiten_ns(const double* gram, double scale_host, int add_identity, int C, double* w_out, double* work,
                              int* info_out, void* stream) {
  if (!gram || !w_out || !work || C < 2 || C > NS_MAX_C || !(scale_host > 0.0)) return WCTB_E_BADARG;
  const size_t smem = (size_t)C * C * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    WCTB_CUDA_TRY(cudaFuncSetAttribute(whiten_ns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)NS_MAX_C * NS_MAX_C * sizeof(double))));
    attr_done = true;
  }
  void* args[] = {(void*)&gram, (void*)&scale_host, (void*)&add_identity, (void*)&C, (void*)&w_out, (void*)&work, (void*)&info_out};
  WCTB_CUDA_TRY(cudaLaunchCooperativeKernel((void*)whiten_ns_kernel, dim3(NS_CTAS), dim3(NS_THREADS), args, smem, (cudaStream_t)stream));
  return WCTB_OK;
}
