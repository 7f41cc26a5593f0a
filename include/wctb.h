/*
 * wctb.h -- C ABI of libwctb.so: the B200 (sm_100a) kernels behind the WCT stylization
 * hot path of MingSun-Tse/Collaborative-Distillation (PytorchWCT/WCT.py + util_wct.py +
 * model/model_{cd,original,kd2sd}.py).
 *
 * The reference has no FFI of its own (it is pure PyTorch); its boundary for this path is
 * the Python call surface listed below.  Each entry point names the reference call it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the Python side.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (e.g. torch Tensor.data_ptr())
 *    unless the name ends in _host;
 *  - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it:
 *    no allocation, no synchronisation, no global state;
 *  - return value: WCTB_OK (0) or a negative WCTB_E_* code; never throws;
 *  - activations between layers use the "P4" layout  [C/4][H][W][4] fp32  (channel-chunk
 *    planar: one float4 = 4 consecutive channels of one pixel).  Images and the public
 *    feature tensors are plain NCHW fp32 (batch 1), converted at the boundary.
 */
#ifndef WCTB_H_
#define WCTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WCTB_ABI_VERSION 1

enum {
  WCTB_OK = 0,
  WCTB_E_BADARG = -1,     /* shape / alignment / enum out of range */
  WCTB_E_UNSUPPORTED = -2,/* configuration not built (e.g. channel count) */
  WCTB_E_WORKSPACE = -3,  /* workspace too small */
  WCTB_E_CUDA = -4        /* a CUDA runtime call failed; see wctb_last_cuda_error() */
};

/* epilogue of a conv layer (what follows conv+bias+ReLU in the reference forward()) */
enum {
  WCTB_EPI_NONE = 0,
  WCTB_EPI_POOL2 = 1, /* nn.MaxPool2d(2,2) floor mode fused on the output (model_cd.py:709,727) */
  WCTB_EPI_UP2 = 2,   /* nn.UpsamplingNearest2d(2) fused on the output   (model_cd.py:261,278) */
  WCTB_EPI_NCHW3 = 3  /* TF32 engine only: last decoder layer run with its 3 output channels zero-padded to 16; the
                         epilogue writes channels 0..2 as NCHW planes [3][H][W] (model_cd.py:84 for stage 1)        */
};

/* which arithmetic a P4->P4 conv uses */
enum {
  WCTB_ENGINE_FP32 = 0, /* CUDA-core FFMA, fp32 exact-order reference path            */
  WCTB_ENGINE_TF32 = 1  /* tcgen05.mma kind::tf32, operands pre-rounded (rna) to TF32 */
};

int wctb_abi_version(void);
const char* wctb_error_string(int code);
int wctb_last_cuda_error(void); /* cudaError_t of the last WCTB_E_CUDA on this thread */

/* ---- layout conversion at the public boundary -------------------------------------------
 * replaces: implicit NCHW tensors flowing between nn.Modules (model_cd.py:724-743).      */
int wctb_nchw_to_p4(const float* src_nchw, float* dst_p4, int C, int H, int W, int round_tf32, void* stream);
int wctb_p4_to_nchw(const float* src_p4, float* dst_nchw, int C, int H, int W, void* stream);

/* ---- weight packing (once, at load time) ---------------------------------------------
 * src: OIHW fp32 [Cout][Cin][3][3] as in the reference state_dict (model_cd.py:690-702).
 * FP32 engine layout:  [tap 9][Cin][Cout] fp32.
 * TF32 engine layout:  [Cin/KG][tap 9][KG/4][Cout][4] fp32 rounded to TF32 (rna); KG = wctb_tf32_kgroup(Cin,Cout) */
int wctb_pack_weights_fp32(const float* w_oihw, float* dst, int Cin, int Cout, void* stream);
int wctb_pack_weights_tf32(const float* w_oihw, float* dst, int Cin, int Cout, void* stream);
int wctb_tf32_kgroup(int Cin, int Cout);
int wctb_tf32_supported(int Cin, int Cout);

/* ---- convolutions: ReflectionPad2d(1) + Conv2d(3x3) + bias + ReLU (+pool / +upsample) ----
 * replaces: `self.relu(self.convXY(self.pad(y)))` (+ `self.pool` / `self.unpool`) in
 *   SmallEncoder{1..5}_16x_aux.forward (model_cd.py:346-349,403-409,485-494,589-603,724-743),
 *   SmallDecoder{1..5}_16x.forward (model_cd.py:83-85,117-122,159-167,211-224,276-294),
 *   Encoder{1..5} and Decoder{1..5} forward (model_original.py:36-39 ... 581-599).
 * round_tf32 != 0: the stored activation is rounded to TF32 (round-to-nearest-away) so that a
 *   following WCTB_ENGINE_TF32 layer reads exactly representable operands.                */

/* first layer: x NCHW [3][H][W] -> y P4 [Cout/4][H][W][4]; conv0 (1x1, model_cd.py:725) is folded
 * into `w` by the host (exact under reflection padding).  w: [tap][3][Cout], Cout % 4 == 0. */
int wctb_conv3x3_first(const float* x_nchw, const float* w, const float* bias, float* y_p4,
                       int H, int W, int Cout, int round_tf32, void* stream);

/* middle layers: P4 -> P4.  Output is [Cout/4][Ho][Wo][4] with (Ho,Wo) = (H,W), (H/2,W/2) or (2H,2W). */
int wctb_conv3x3_p4(const float* x_p4, const float* w_packed, const float* bias, float* y_p4,
                    int H, int W, int Cin, int Cout, int epilogue, int round_tf32, int engine,
                    void* stream);

/* fused encoder head (TF32 engine): conv11 (3 -> C1, FFMA, conv0 folded) + ReLU feeding conv12 (C1 -> Cout, tcgen05)
 * + ReLU (+pool) in one kernel; the C1-channel full-resolution activation stays in shared memory.
 * replaces: y = relu(conv11(pad(conv0(y)))); y = relu(conv12(pad(y))); [y = pool(y)]   (model_cd.py:725-728)
 * w11: [tap][3][C1] fp32, w12_packed: wctb_pack_weights_tf32 layout.  Supported (C1,Cout): (16,16), (64,64). */
int wctb_conv_head_supported(int C1, int Cout);
int wctb_conv_head(const float* x_nchw, const float* w11, const float* b11, const float* w12_packed,
                   const float* b12, float* y_p4, int H, int W, int C1, int Cout, int epilogue,
                   int round_tf32, void* stream);

/* all-tensor-core variant of the fused head for the 16x nets (C1 = Cout = 16): conv11 runs on tcgen05 as well.
 * w11_tc: [3 dy][2][2 chunks][16][4] fp32 rounded to TF32, chunk c of MMA (dy,h) holding tap dx = 2h + c (RGB + a zero
 * channel; dx = 3 is all zero) -- built by the host from the conv0-folded conv11 weights (see nets.py).            */
int wctb_conv_head_tc(const float* x_nchw, const float* w11_tc, const float* b11, const float* w12_packed,
                      const float* b12, float* y_p4, int H, int W, int epilogue, int round_tf32, void* stream);

/* fused decoder tail (TF32 engine): conv12 (16 -> 16) + ReLU + conv11 (16 -> 3) + ReLU in one kernel,
 * output NCHW [3][H][W]; the 16-channel full-resolution intermediate stays in shared memory.  With
 * upsample_input != 0, x_p4 is the HALF-resolution tensor [Cin/4][H/2][W/2][4] and nn.UpsamplingNearest2d(2) is applied
 * while the operand tile is filled, so the upsampled tensor never exists in HBM.
 * replaces: [y = unpool(y);] y = relu(conv12(pad(y))); y = relu(conv11(pad(y)))      (model_cd.py:291-293)
 * Both convs run on tcgen05.  w12_packed: wctb_pack_weights_tf32 layout; w11: conv11 OIHW weights zero-padded to 16 output
 * channels and packed with wctb_pack_weights_tf32 (Cin = Cmid, Cout = 16); b11: [3].  Supported (Cin,Cmid): (16,16).   */
int wctb_conv_tail_supported(int Cin, int Cmid);
int wctb_conv_tail(const float* x_p4, const float* w12_packed, const float* b12, const float* w11,
                   const float* b11, float* y_nchw, int H, int W, int Cin, int Cmid, int upsample_input,
                   void* stream);

/* last decoder layer: P4 [Cin/4][H][W][4] -> NCHW [3][H][W], ReLU kept (model_cd.py:293). w: [tap][Cin][3] */
int wctb_conv3x3_last(const float* x_p4, const float* w, const float* bias, float* y_nchw,
                      int H, int W, int Cin, void* stream);

/* ---- "h2" engine: fp32-accurate tensor-core convolutions (csrc/conv_h2.cu) -----------------------------------
 * replaces the same reference lines as wctb_conv3x3_p4 above (`self.relu(self.convXY(self.pad(y)))` [+ pool / unpool],
 * model_cd.py:724-743, 276-294; model_original.py:492-511, 581-599) at fp32 accuracy ON the tensor cores: every fp32
 * operand is carried as a pair of fp16 numbers (x = hi + lo, 22 significand bits) and a product is evaluated as
 * hi*w_hi + lo*w_hi + hi*w_lo by tcgen05.mma.kind::f16 with fp32 accumulation.
 * Activation layout H8: [ceil(C/8)][2 (hi, lo)][H][W][8] fp16 (same bytes as fp32).
 * Packed weights: [Cout/N][ceil(Cin/16)][tap 9][2 k-chunks][hi N rows | lo N rows][8] fp16 of w * s, with s a power of two
 * chosen ON THE DEVICE from max|w| (no host sync); wscale is a caller-owned device buffer of 4 floats:
 * [0] scratch, [1] = 1/s (read by the conv kernel), [2] = s.  N = Cout for Cout in {16,32,64,128}, else 128.
 * wctb_conv3x3_h2 writes y_h8 (H8, for the next h2 layer) and / or y_p4 (fp32 P4, for the statistics kernels and the
 * public NCHW API); either may be NULL.  epilogue WCTB_EPI_NCHW3: Cout must be 16 (3 real channels zero-padded by the
 * caller), y_p4 is the [3][H][W] image.  Supported: Cin % 8 == 0, Cout in {16,32,64,128} or a multiple of 128.          */
int wctb_h2_supported(int Cin, int Cout);
long long wctb_h2_packed_halves(int Cin, int Cout);
int wctb_pack_weights_h2(const float* w_oihw, void* dst_halves, float* wscale, int Cin, int Cout, void* stream);
int wctb_conv3x3_h2(const void* x_h8, const void* w_packed, const float* bias, const float* wscale, void* y_h8,
                    float* y_p4, int H, int W, int Cin, int Cout, int epilogue, void* stream);
/* first layer for the h2 engine: x NCHW [3][H][W] -> H8 and / or fp32 P4, fp32 FFMA (conv0 folded by the host),
 * same arithmetic as wctb_conv3x3_first.  w: [tap][3][Cout] (wctb_pack_weights_fp32), Cout % 8 == 0.                   */
int wctb_conv3x3_first_h2(const float* x_nchw, const float* w, const float* bias, void* y_h8, float* y_p4,
                          int H, int W, int Cout, void* stream);
/* fused encoder head of the 16x nets on the h2 engine (csrc/conv_h2_fused.cu): conv11 (3 -> 16, conv0 folded) + ReLU +
 * conv12 (16 -> 16) + ReLU + MaxPool2d(2,2) in one persistent kernel, both convs on tcgen05 with the three horizontal
 * filter taps stacked along N; the 16-channel full-resolution activations never touch HBM.
 * replaces: y = relu(conv11(pad(conv0(y)))); y = relu(conv12(pad(y))); y = pool(y)      (model_cd.py:725-728)
 * w11_packed: [2][2 k-chunks][96][8] fp16, w12_packed: [3 dy][2 k-chunks][96][8] fp16 (rows dx*16+co: hi, 48+dx*16+co: lo)
 * of w * s with a host-chosen power-of-two s; inv_s = 1/s.  Built by ops.pack_head_h2_w11 / ops.pack_dx_h2.  y: H8 16 ch. */
int wctb_conv_head_h2(const float* x_nchw, const void* w11_packed, const float* b11, float inv_s11,
                      const void* w12_packed, const float* b12, float inv_s12, void* y_h8, int H, int W, void* stream);
/* fused decoder tail of the 16x nets on the h2 engine: [UpsamplingNearest2d(2) +] conv12 (16 -> 16) + ReLU + conv11 (16 -> 3)
 * + ReLU -> NCHW fp32 image, same kernel structure as the head (dx-stacked taps, block-pipelined TMEM rings); with
 * upsample_input != 0 x_h8 is the HALF-resolution tensor and the nearest x2 is applied while the operand tile is loaded.
 * replaces: [y = unpool(y);] y = relu(conv12(pad(y))); y = relu(conv11(pad(y)))       (model_cd.py:291-293)
 * w12_packed / w11_packed: ops.pack_dx_h2 layout (conv11's 3 output channels in rows dx*16 + {0,1,2}); b11: [3].        */
int wctb_conv_tail_h2(const void* x_h8, const void* w12_packed, const float* b12, float inv_s12, const void* w11_packed,
                      const float* b11, float inv_s11, float* y_nchw, int H, int W, int upsample_input, void* stream);
/* Strip-sharded variant of the fused tail (multi-GPU): compute + halo exchange in ONE kernel.  The image is
 * computed for the whole extended strip (W columns), but only the rank's own columns [own_x0, own_x0 + own_w) are kept; they
 * are written to `out` -- the rank's NEXT-stage extended strip [3][H][out_pitch] at column out_x0 -- and the `halo` columns next
 * to a seam are additionally stored straight into the neighbours' next-stage strips through peer-mapped pointers (st.global
 * over NVLink: peer_l = left neighbour's buffer, our columns land at peer_l_x0; peer_r likewise; NULL at a true image border).
 * No pack / send / recv / unpack: the next stage starts after one cross-rank barrier.  All pointers are device pointers valid
 * in THIS process (peer buffers opened through CUDA IPC / symmetric memory by the caller).                                  */
typedef struct wctb_tail_shard {
  float* out;    int out_pitch, out_x0;
  int own_x0, own_w, halo;
  float* peer_l; int peer_l_pitch, peer_l_x0;
  float* peer_r; int peer_r_pitch, peer_r_x0;
} wctb_tail_shard;
int wctb_conv_tail_h2_sharded(const void* x_h8, const void* w12_packed, const float* b12, float inv_s12, const void* w11_packed,
                              const float* b11, float inv_s11, int H, int W, int upsample_input, const wctb_tail_shard* shard,
                              void* stream);
/* layout conversion: NCHW fp32 <-> H8, fp32 P4 -> H8 */
int wctb_nchw_to_h8(const float* src_nchw, void* dst_h8, int C, int H, int W, void* stream);
int wctb_h8_to_nchw(const void* src_h8, float* dst_nchw, int C, int H, int W, void* stream);
int wctb_p4_to_h8(const float* src_p4, void* dst_h8, int C, int H, int W, void* stream);

/* ---- WCT statistics ------------------------------------------------------------------
 * replaces: torch.mean(cF,1) / cF - mean / torch.mm(cF, cF.t()) (util_wct.py:68-70, 94-96).
 * x is P4 [C/4][H][W][4]; the sums run over the region rows [y0,y1) x cols [x0,x1) only
 * (the whole map for one GPU; the rank's own strip without halo when sharded).
 * channel_sum:   sum_out[C]  (fp64)  += sum over region of x          (caller zeroes it)
 * centered_gram: gram_out[C*C] (fp64, row-major, full symmetric) += sum (x-mean)(x-mean)^T  */
int wctb_channel_sum(const float* x_p4, int C, int H, int W, int y0, int y1, int x0, int x1,
                     double* sum_out, void* stream);
int wctb_centered_gram(const float* x_p4, int C, int H, int W, int y0, int y1, int x0, int x1,
                       const double* mean, double* gram_out, void* stream);
/* same contract; products and per-stage partial sums in fp32 (flushed to fp64 every 128 pixels): Gram accurate to ~1e-7
 * relative -- used with the TF32 conv engine, whose feature noise (1e-3) is far larger.                              */
int wctb_centered_gram_fast(const float* x_p4, int C, int H, int W, int y0, int y1, int x0, int x1,
                            const double* mean, double* gram_out, void* stream);

/* ---- symmetric eigendecomposition (one-sided Jacobi on the pivoted Cholesky factor, fp64) ---
 * replaces: torch.svd(contentConv, some=False) / torch.svd(styleConv) (util_wct.py:74,100);
 * only (E, V) are consumed there and the matrices are symmetric PSD.
 * a: nprob (<= 8) matrices [C][C] fp64 (row-major symmetric), each scaled by scale_host[prob] (HOST array) and, if
 * add_identity, + I (the `--numpy` variant, util_wct.py:143) before the solve.
 * Outputs per problem: evals[C] (>= 0, unsorted), evecs[C][C] column k = unit eigenvector k
 * stored as evecs[k*C + i] (zero vector when the eigenvalue is exactly 0, and -- for C <= 128 -- for the null space
 * beyond the numerical rank, i.e. eigenvalues below ~1e-14 of the largest diagonal entry, which are reported as 0).
 * work: nprob*C*C + 16 doubles of scratch.  sweeps_out (optional, may be NULL): int[nprob].    */
int wctb_eigh_jacobi(const double* a, int nprob, int C, const double* scale_host, int add_identity,
                     double* evals, double* evecs, double* work, int* sweeps_out, void* stream);

/* same with an explicit early-stop threshold for the C <= 128 solver: the iteration ends after a sweep whose largest
 * |cos(column_p, column_q)| is below early_stop_cos (quadratic convergence: that sweep leaves ~early_stop_cos^2).
 * wctb_eigh_jacobi uses 3e-6 (residual <= 1e-9); 1e-4 still leaves <= ~1e-8 and usually saves one sweep; 1e-2 leaves
 * <= ~1e-5 (whitening matrix error ~2e-6) and saves two -- enough when the features carry TF32 noise (1e-3).
 * Range [0, 0.1]; ignored by the C > 128 solver.                                                                       */
int wctb_eigh_jacobi_tol(const double* a, int nprob, int C, const double* scale_host, int add_identity,
                         double early_stop_cos, double* evals, double* evecs, double* work, int* sweeps_out,
                         void* stream);

/* ---- whitening / colouring matrix ----------------------------------------------------
 * replaces: util_wct.py:117-126 + the alpha blend of transform() (util_wct.py:219):
 *   W   = sum_{k: Ec_k > tau*max(Ec)} Ec_k^-1/2 vc_k vc_k^T        (117-119)
 *   Col = sum_{k: Es_k > tau*max(Es)} Es_k^+1/2 vs_k vs_k^T        (124-125)
 *   M   = alpha * Col W + (1-alpha) I ;  b = alpha*mean_s + (1-alpha)*mean_c
 * so that  csF = M (cF - mean_c) + b.   (tau replaces EigenValueThre=1e-100)
 * Outputs fp32: m_out [C][C] row-major (row = output channel), b_out [C], mean_c_out [C].
 * work: 3*C*C + 8 doubles.                                                                  */
int wctb_wct_matrix(const double* c_evals, const double* c_evecs, const double* c_mean,
                    const double* s_evals, const double* s_evecs, const double* s_mean,
                    int C, double tau, double alpha, float* m_out, float* b_out,
                    float* mean_c_out, double* work, void* stream);

/* same, with the reference's eigenvalue-truncation knobs (util_wct.py:26-27 NumEigenValue / RatEigenValue; their uses at
 * :87-88 and :113-114 are commented out in the reference): only the keep_c (content) / keep_s (style) LARGEST directions
 * are used, and of those only the ones above tau*max as before; keep <= 0 or >= C keeps all (== wctb_wct_matrix).      */
int wctb_wct_matrix_topk(const double* c_evals, const double* c_evecs, const double* c_mean,
                         const double* s_evals, const double* s_evecs, const double* s_mean,
                         int C, double tau, double alpha, int keep_c, int keep_s, float* m_out, float* b_out,
                         float* mean_c_out, double* work, void* stream);

/* ---- whitening matrix without an eigendecomposition (opt-in; not yet run on hardware, see csrc/whiten_ns.cu) --------
 * W = (scale*gram [+ I])^-1/2, pseudo-inverse on the range (rank-revealing pivoted Cholesky + coupled Newton-Schulz on
 * L^T L, GEMMs only, cooperative grid of 16 CTAs): the content half of util_wct.py:74,117-119 in one launch.
 * gram: fp64 [C][C] centred Gram (as produced by wctb_centered_gram*), C <= 128; w_out: fp64 [C][C] symmetric, zero
 * rows/columns for dead channels; work: wctb_workspace_doubles(WCTB_WS_WHITEN_NS, C, 1) doubles;
 * info_out (optional): int[3] = {rank, iterations, converged}.                                                         */
int wctb_whiten_ns(const double* gram, double scale_host, int add_identity, int C, double* w_out, double* work,
                   int* info_out, void* stream);
/* M, b, mean_c from a ready whitening matrix W (wctb_whiten_ns) and the style eigensystem: M = alpha*Col*W + (1-alpha) I.
 * work: wctb_workspace_doubles(WCTB_WS_WCT_MATRIX, C, 1) doubles.                                                      */
int wctb_wct_matrix_w(const double* w_whiten, const double* c_mean, const double* s_evals, const double* s_evecs,
                      const double* s_mean, int C, double tau, double alpha, float* m_out, float* b_out,
                      float* mean_c_out, double* work, void* stream);

/* csF = M (cF - mean_c) + b on a P4 map of npix pixels (whole extended strip).
 * replaces: torch.mm(step2, cF), torch.mm(..., whiten_cF), + s_mean (util_wct.py:120,125,126). */
int wctb_wct_apply(const float* x_p4, const float* m, const float* b, const float* mean_c,
                   float* y_p4, int C, long long npix, int round_tf32, void* stream);

/* Fold csF = M(x-mean_c)+b into the decoder's first conv (exact: the conv is linear before its
 * ReLU and reflection padding maps constants to constants):
 *   w_out[o][i][t] = sum_j w_oihw[o][j][t] * M[j][i]
 *   b_out[o]       = bias[o] + sum_{j,t} w_oihw[o][j][t] * (b[j] - (M mean_c)[j])
 * w_out is OIHW fp32 (pack it afterwards).                                                */
int wctb_fold_wct_into_conv(const float* w_oihw, const float* bias, const float* m, const float* b,
                            const float* mean_c, float* w_out, float* b_out, int Cin, int Cout,
                            void* stream);

/* ---- workspace sizes -----------------------------------------------------------------------------------------
 * HOST query (no GPU): doubles of scratch the `work` argument of an entry point must hold, so that a caller can size one
 * persistent arena instead of allocating per call (the reference frees and re-allocates through empty_cache(), WCT.py:99-105). */
enum { WCTB_WS_EIGH = 0 /* wctb_eigh_jacobi[_tol]: nprob*C*C + 16 */, WCTB_WS_WCT_MATRIX = 1 /* wctb_wct_matrix[_topk|_w]: 3*C*C + 8 */,
       WCTB_WS_WHITEN_NS = 2 /* wctb_whiten_ns: 8*C*C + 8 */ };
long long wctb_workspace_doubles(int op, int C, int nprob);

/* ---- strip halos of the multi-GPU path (no counterpart in the single-GPU reference) ------------------
 * An image strip is NCHW [C][H][W] fp32.  pack: columns [x0, x0+w) -> contiguous send buffer [C][H][w].
 * unpack: buffer [C][H][w] -> columns [x0, x0+w) of the extended strip [C][H][We] (also used to place the rank's own
 * columns).  Plain copies (bit-exact); the buffers are what the host hands to NCCL send/recv.                          */
int wctb_halo_pack(const float* img_nchw, float* buf, int C, int H, int W, int x0, int w, void* stream);
int wctb_halo_unpack(const float* buf, float* ext_nchw, int C, int H, int We, int x0, int w, void* stream);

/* ---- image I/O around the path: byte / integer kernels, bit-exact --------------------
 * Images here are interleaved 8-bit RGB, [H][W][3] (what PIL / nvJPEG produce and consume).
 *
 * u8hwc_to_nchw  replaces transforms.ToTensor() (reference data_loader.py:56-57): fp32 [3][H][W] = u8 / 255
 *                (correctly rounded fp32 division, as torch's `.to(float32).div(255)`).
 * nchw_to_u8hwc  replaces the quantisation inside vutils.save_image() (reference WCT.py:128):
 *                u8 = trunc(clamp(x*255 + 0.5, 0, 255)) with the two fp32 roundings of `mul(255).add_(0.5)`.        */
int wctb_u8hwc_to_nchw(const uint8_t* src_hwc, float* dst_nchw, int H, int W, void* stream);
int wctb_nchw_to_u8hwc(const float* src_nchw, uint8_t* dst_hwc, int H, int W, void* stream);

/* 8-bit antialiased bilinear resize = transforms.Resize(size) on a PIL image (reference data_loader.py:52-55), i.e.
 * Pillow's ImagingResample for 8-bit pixels: per output pixel a normalised triangle filter of support
 * max(in/out, 1), coefficients in 22-bit fixed point, horizontal pass then vertical pass through an 8-bit intermediate.
 *   wctb_resize_ksize        HOST: taps per output pixel for in_size -> out_size (>= 3), or WCTB_E_BADARG
 *   wctb_resize_coeffs_host  HOST (needs no GPU): bounds_host[2*out] = (first tap, tap count),
 *                            coeffs_host[out*ksize] = fixed-point weights; the caller copies both to the device
 *   wctb_resize_u8_pass      DEVICE: resample axis 1 (width: dst is [H][out][3]) or axis 0 (height: dst is [out][W][3])
 * A full resize is: pass(axis=1) if the width changes, then pass(axis=0) if the height changes (PIL's order).       */
int wctb_resize_ksize(int in_size, int out_size);
int wctb_resize_coeffs_host(int in_size, int out_size, int* bounds_host, int* coeffs_host);
int wctb_resize_u8_pass(const uint8_t* src_hwc, uint8_t* dst_hwc, int H, int W, int out_size, int axis,
                        const int* bounds, const int* coeffs, int ksize, void* stream);

/* debug: 0 (default) = Cholesky-preconditioned Jacobi for C <= 128; 1 = legacy Jacobi on the matrix itself (A/B timing,
 * tools/eig_diag.py).  Process-wide; not meant to be toggled while work is in flight.                                */
int wctb_debug_set_eigh_variant(int variant);

/* debug: which kernel wctb_centered_gram_fast uses for C = 24 / 32 (A/B timing, tools/gram_ab.py):
 * 0 (default) = register-resident accumulation (a thread owns whole pixels) fed from a cp.async shared-memory ring;
 * 2 = the same accumulation fed by direct global loads through L1; 1 = the staged shared-memory tile kernel everywhere;
 * 3 = variant 0 with the range-checked last iteration peeled out of the loop; 4 = two pixels per thread and iteration,
 * peeled (3 and 4 are not yet run on hardware).                                                                       */
int wctb_debug_set_gram_variant(int variant);

/* debug: 0 (default) = wctb_conv3x3_first computes two pixels per thread when W >= 64; 1 = one pixel per thread (A/B;
 * the two kernels are bit-identical).                                                                                 */
int wctb_debug_set_first_variant(int variant);

/* debug: when buf16 != NULL the C in (64,128] solve records clock64() phase times of thread 0 into buf16:
 * [0] load + compaction, [1] Cholesky, [2] Jacobi sweeps, [8] sweep count, [9] live size k.                          */
int wctb_debug_eigh_profile(long long* buf16);
/* debug: fp64 pipe probe -- out3[0] = cycles of a 4096-long dependent DFMA chain, out3[1] = cycles for a 512-thread CTA
 * to issue 4096 DFMAs per thread in 8 independent chains.                                                           */
int wctb_debug_dp_rate(long long* out3, void* stream);

/* debug: when buf != NULL a few CTAs of the fused head record clock64() phase stamps into buf[128] (tools/trace_head.py) */
int wctb_debug_set_trace(long long* buf);

/* debug: cycles for iters*4*nacc tcgen05.mma (M=128,K=8 tf32) per CTA; layout 0 = planes (SWIZZLE_NONE), 1 = SWIZZLE_128B, 2 = LBO 16 */
int wctb_debug_mma_rate(long long* out_cycles, int N, int layout, int nacc, int iters, int ctas, void* stream);
/* cycles per tcgen05.mma.kind::f16 (M=128, K=16) vs N / number of independent accumulators; tcgen05.ld throughput */
/* debug: 1 (default) = h2 conv layers with Cout <= 64 and Cin <= 64 keep their packed weights resident in shared memory (and the
   64-channel ones stack the weight halves along N); 0 = weights streamed with every pipeline stage (A/B timing) */
int wctb_debug_set_h2_resident(int on);
int wctb_debug_mma_rate_f16(long long* out_cycles, int N, int nacc, int iters, int ctas, void* stream);
/* same, the A descriptor of MMA k starts (k % 3) * a_off16 sixteen-byte units past a 1024-byte boundary (the dx taps of the
   linear-pitch convolution: does a start address off the 128-byte core-matrix grid slow the operand fetch?) */
int wctb_debug_mma_rate_f16_off(long long* out_cycles, int N, int nacc, int iters, int ctas, int a_off16, void* stream);
int wctb_debug_ldtm_rate(long long* out_cycles, int nwarps, int per_iter, int iters, int ctas, void* stream);

/* ---- self tests (device-side descriptor / pipeline checks used by tests and smoke) ------- */
int wctb_selftest_umma(float* out_128xN, const float* a_128xK, const float* b_NxK, int N, int K,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WCTB_H_ */
