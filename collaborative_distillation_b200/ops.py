"""Tensor-level wrappers over the C ABI (include/wctb.h).  torch is used for device memory and
streams only; all arithmetic on the path happens inside libwctb.so.  CUDA tensors only."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_NONE, EPI_POOL2, EPI_UP2, EPI_NCHW3, ENGINE_FP32, ENGINE_TF32, ENGINE_H2, check  # noqa: F401

_launches = 0  # number of libwctb kernel-launching calls (bench.py reports kernels via its own table)
KERNELS_PER_CALL = {"whiten_ns": 1, "wct_matrix_w": 3, "halo": 1, "nchw_to_p4": 1, "p4_to_nchw": 1, "pack_fp32": 1, "pack_tf32": 1, "conv_first": 1, "conv_p4": 1,
                    "conv_last": 1, "conv_head": 1, "conv_head_tc": 1, "conv_tail": 1, "channel_sum": 1, "centered_gram": 1, "eigh": 1, "wct_matrix": 5, "wct_apply": 1,
                    "fold": 2, "pack_h2": 2, "conv_h2": 1, "conv_first_h2": 1, "to_h8": 1, "from_h8": 1, "conv_head_h2": 1, "conv_tail_h2": 1}


def launches() -> int:
    return _launches


def add_launches(n: int):
    """a replayed CUDA graph re-issues the kernels recorded at capture time"""
    global _launches
    _launches += int(n)


def _count(kind):
    global _launches
    _launches += KERNELS_PER_CALL[kind]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need(t: torch.Tensor, dtype=torch.float32):
    if not t.is_cuda:
        raise _lib.WctbError("libwctb ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise _lib.WctbError("expected contiguous %s tensor, got %s contiguous=%s" % (dtype, t.dtype, t.is_contiguous()))
    return t.data_ptr()


def nchw_to_p4(x: torch.Tensor, round_tf32: bool = False) -> torch.Tensor:
    """[C,H,W] (or [1,C,H,W]) -> P4 [C/4,H,W,4]"""
    if x.dim() == 4:
        x = x.squeeze(0)
    C, H, W = x.shape
    y = torch.empty(C // 4, H, W, 4, device=x.device, dtype=torch.float32)
    check(_lib.load().wctb_nchw_to_p4(_need(x), _need(y), C, H, W, int(round_tf32), _stream()), "nchw_to_p4")
    _count("nchw_to_p4")
    return y


def p4_to_nchw(x: torch.Tensor) -> torch.Tensor:
    """P4 [C/4,H,W,4] -> [1,C,H,W]"""
    C4, H, W, _ = x.shape
    y = torch.empty(1, C4 * 4, H, W, device=x.device, dtype=torch.float32)
    check(_lib.load().wctb_p4_to_nchw(_need(x), _need(y), C4 * 4, H, W, _stream()), "p4_to_nchw")
    _count("p4_to_nchw")
    return y


def halo_pack(img: torch.Tensor, x0: int, w: int) -> torch.Tensor:
    """columns [x0, x0+w) of an NCHW strip [1,C,H,W] -> contiguous [1,C,H,w] (the buffer a rank sends to a neighbour)"""
    _, C, H, W = img.shape
    buf = torch.empty(1, C, H, w, device=img.device, dtype=torch.float32)
    check(_lib.load().wctb_halo_pack(_need(img), _need(buf), C, H, W, int(x0), int(w), _stream()), "halo_pack")
    _count("halo")
    return buf


def halo_unpack(buf: torch.Tensor, ext: torch.Tensor, x0: int) -> None:
    """write a contiguous [1,C,H,w] buffer into columns [x0, x0+w) of the extended strip ext [1,C,H,We]"""
    _, C, H, w = buf.shape
    We = ext.shape[-1]
    check(_lib.load().wctb_halo_unpack(_need(buf), _need(ext), C, H, We, int(x0), int(w), _stream()), "halo_unpack")
    _count("halo")


def tf32_supported(cin: int, cout: int) -> bool:
    return bool(_lib.load().wctb_tf32_supported(cin, cout))


def pack_weights(w_oihw: torch.Tensor, engine: int) -> torch.Tensor:
    cout, cin = w_oihw.shape[:2]
    w = w_oihw.detach().contiguous().float()
    dst = torch.empty(9 * cin * cout, device=w.device, dtype=torch.float32)
    lib = _lib.load()
    if engine == ENGINE_TF32:
        check(lib.wctb_pack_weights_tf32(_need(w), _need(dst), cin, cout, _stream()), "pack_weights_tf32")
        _count("pack_tf32")
    else:
        check(lib.wctb_pack_weights_fp32(_need(w), _need(dst), cin, cout, _stream()), "pack_weights_fp32")
        _count("pack_fp32")
    return dst


def conv3x3_first(x_nchw: torch.Tensor, w: torch.Tensor, b: torch.Tensor, cout: int, round_tf32: bool) -> torch.Tensor:
    """x: [1,3,H,W] or [3,H,W]; w packed [9][3][cout] -> P4 [cout/4,H,W,4]"""
    if x_nchw.dim() == 4:
        x_nchw = x_nchw.squeeze(0)
    _, H, W = x_nchw.shape
    y = torch.empty(cout // 4, H, W, 4, device=x_nchw.device, dtype=torch.float32)
    check(_lib.load().wctb_conv3x3_first(_need(x_nchw), _need(w), _need(b), _need(y), H, W, cout, int(round_tf32), _stream()),
          "conv3x3_first")
    _count("conv_first")
    return y


def conv3x3_p4(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, cout: int, epilogue: int, round_tf32: bool,
               engine: int) -> torch.Tensor:
    C4, H, W, _ = x.shape
    if epilogue == EPI_POOL2:
        Ho, Wo = H // 2, W // 2
    elif epilogue == EPI_UP2:
        Ho, Wo = 2 * H, 2 * W
    else:
        Ho, Wo = H, W
    if epilogue == EPI_NCHW3:     # last decoder layer on the tensor cores: cout is the padded 16, output is the NCHW image
        y = torch.empty(1, 3, H, W, device=x.device, dtype=torch.float32)
    else:
        y = torch.empty(cout // 4, Ho, Wo, 4, device=x.device, dtype=torch.float32)
    check(_lib.load().wctb_conv3x3_p4(_need(x), _need(w), _need(b), _need(y), H, W, C4 * 4, cout, epilogue,
                                      int(round_tf32), engine, _stream()), "conv3x3_p4")
    _count("conv_p4")
    return y


def conv_head_supported(c1: int, cout: int) -> bool:
    return bool(_lib.load().wctb_conv_head_supported(c1, cout))


def conv_head(x_nchw, w11, b11, w12_packed, b12, c1: int, cout: int, epilogue: int, round_tf32: bool) -> torch.Tensor:
    """fused conv11(3->c1)+ReLU+conv12(c1->cout)+ReLU(+pool): image [1,3,H,W] -> P4 [cout/4,Ho,Wo,4]"""
    if x_nchw.dim() == 4:
        x_nchw = x_nchw.squeeze(0)
    _, H, W = x_nchw.shape
    Ho, Wo = (H // 2, W // 2) if epilogue == EPI_POOL2 else (H, W)
    y = torch.empty(cout // 4, Ho, Wo, 4, device=x_nchw.device, dtype=torch.float32)
    check(_lib.load().wctb_conv_head(_need(x_nchw), _need(w11), _need(b11), _need(w12_packed), _need(b12), _need(y), H, W,
                                     c1, cout, epilogue, int(round_tf32), _stream()), "conv_head")
    _count("conv_head")
    return y


def tf32_round(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest (ties away) to TF32, for host-side weight preparation"""
    u = t.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_head_tc_weights(w_oihw: torch.Tensor) -> torch.Tensor:
    """conv11 [16][3][3][3] (conv0 folded) -> [3 dy][2 h][2 c][16 n][4 e] TF32: chunk c of MMA (dy,h) = tap dx = 2h+c"""
    t = torch.zeros(3, 2, 2, 16, 4, device=w_oihw.device, dtype=torch.float32)
    for dy in range(3):
        for hh in range(2):
            for c in range(2):
                dx = 2 * hh + c
                if dx <= 2:
                    t[dy, hh, c, :, :3] = w_oihw[:, :, dy, dx]
    return tf32_round(t).contiguous()


def conv_head_tc(x_nchw, w11_tc, b11, w12_packed, b12, epilogue: int, round_tf32: bool) -> torch.Tensor:
    """all-tensor-core fused head of the 16x nets: image [1,3,H,W] -> P4 [4,Ho,Wo,4]"""
    if x_nchw.dim() == 4:
        x_nchw = x_nchw.squeeze(0)
    _, H, W = x_nchw.shape
    Ho, Wo = (H // 2, W // 2) if epilogue == EPI_POOL2 else (H, W)
    y = torch.empty(4, Ho, Wo, 4, device=x_nchw.device, dtype=torch.float32)
    check(_lib.load().wctb_conv_head_tc(_need(x_nchw), _need(w11_tc), _need(b11), _need(w12_packed), _need(b12), _need(y),
                                        H, W, epilogue, int(round_tf32), _stream()), "conv_head_tc")
    _count("conv_head_tc")
    return y


def conv_tail_supported(cin: int, cmid: int) -> bool:
    return bool(_lib.load().wctb_conv_tail_supported(cin, cmid))


def conv_tail(x_p4, w12_packed, b12, w11, b11, upsample_input: bool) -> torch.Tensor:
    """fused [nearest x2 +] conv12(16->16)+ReLU+conv11(16->3)+ReLU: P4 [4,h,w,4] -> image [1,3,H,W].
    w11: conv11 weights zero-padded to 16 outputs and packed for the TF32 engine; b11: [3]."""
    C4, h, w, _ = x_p4.shape
    H, W = (2 * h, 2 * w) if upsample_input else (h, w)
    y = torch.empty(1, 3, H, W, device=x_p4.device, dtype=torch.float32)
    check(_lib.load().wctb_conv_tail(_need(x_p4), _need(w12_packed), _need(b12), _need(w11), _need(b11), _need(y), H, W,
                                     C4 * 4, b12.numel(), int(upsample_input), _stream()), "conv_tail")
    assert w11.numel() == 9 * 16 * 16
    _count("conv_tail")
    return y


def conv3x3_last(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    C4, H, W, _ = x.shape
    y = torch.empty(1, 3, H, W, device=x.device, dtype=torch.float32)
    check(_lib.load().wctb_conv3x3_last(_need(x), _need(w), _need(b), _need(y), H, W, C4 * 4, _stream()), "conv3x3_last")
    _count("conv_last")
    return y


def channel_sum(x: torch.Tensor, region=None) -> torch.Tensor:
    """P4 map -> fp64 [C] sums over region (y0,y1,x0,x1) (default: whole map)"""
    C4, H, W, _ = x.shape
    y0, y1, x0, x1 = region if region is not None else (0, H, 0, W)
    out = torch.zeros(C4 * 4, device=x.device, dtype=torch.float64)
    check(_lib.load().wctb_channel_sum(_need(x), C4 * 4, H, W, y0, y1, x0, x1, _need(out, torch.float64), _stream()),
          "channel_sum")
    _count("channel_sum")
    return out


def centered_gram(x: torch.Tensor, mean: torch.Tensor, region=None, out: torch.Tensor = None, fast: bool = False) -> torch.Tensor:
    """P4 map, fp64 mean [C] -> fp64 [C,C] sum (x-mean)(x-mean)^T over the region (accumulated into `out` if given)"""
    C4, H, W, _ = x.shape
    C = C4 * 4
    y0, y1, x0, x1 = region if region is not None else (0, H, 0, W)
    if out is None:
        out = torch.zeros(C, C, device=x.device, dtype=torch.float64)
    fn = _lib.load().wctb_centered_gram_fast if fast else _lib.load().wctb_centered_gram
    check(fn(_need(x), C, H, W, y0, y1, x0, x1, _need(mean, torch.float64), _need(out, torch.float64), _stream()),
          "centered_gram")
    _count("centered_gram")
    return out


def eigh_jacobi(a: torch.Tensor, scale, add_identity: bool = False, return_sweeps: bool = False, early_stop: float = None):
    """a: fp64 [nprob,C,C] symmetric PSD; scale: nprob host floats.  -> evals [nprob,C], evecs [nprob,C(k),C(i)].
    early_stop: |cos| threshold that ends the C <= 128 iteration (None = the library default 3e-6, see wctb.h)."""
    import ctypes
    nprob, C, _ = a.shape
    scale = [float(v) for v in (scale.tolist() if torch.is_tensor(scale) else scale)]
    assert len(scale) == nprob
    scale_host = (ctypes.c_double * nprob)(*scale)
    evals = torch.empty(nprob, C, device=a.device, dtype=torch.float64)
    evecs = torch.empty(nprob, C, C, device=a.device, dtype=torch.float64)
    work = torch.empty(_lib.load().wctb_workspace_doubles(_lib.WS_EIGH, C, nprob), device=a.device, dtype=torch.float64)
    sweeps = torch.zeros(nprob, device=a.device, dtype=torch.int32)
    if early_stop is None:
        check(_lib.load().wctb_eigh_jacobi(_need(a, torch.float64), nprob, C, scale_host, int(add_identity),
                                           _need(evals, torch.float64), _need(evecs, torch.float64),
                                           _need(work, torch.float64), _need(sweeps, torch.int32), _stream()), "eigh_jacobi")
    else:
        check(_lib.load().wctb_eigh_jacobi_tol(_need(a, torch.float64), nprob, C, scale_host, int(add_identity), float(early_stop),
                                               _need(evals, torch.float64), _need(evecs, torch.float64),
                                               _need(work, torch.float64), _need(sweeps, torch.int32), _stream()), "eigh_jacobi_tol")
    _count("eigh")
    return (evals, evecs, sweeps) if return_sweeps else (evals, evecs)


def set_eigh_variant(v: int):
    """debug: 0 = Cholesky-preconditioned Jacobi (default), 1 = legacy Jacobi on the matrix itself (C <= 128 only)"""
    check(_lib.load().wctb_debug_set_eigh_variant(int(v)), "debug_set_eigh_variant")


def set_first_variant(v: int):
    """debug: 0 = conv3x3_first computes two pixels per thread (default), 1 = one pixel per thread (bit-identical)"""
    check(_lib.load().wctb_debug_set_first_variant(int(v)), "debug_set_first_variant")


def set_gram_variant(v: int):
    """debug: fast Gram for C = 24 / 32: 0 = register accumulation fed from a cp.async ring (default), 2 = fed through L1,
    1 = staged shared-memory kernel everywhere, 3 / 4 = later ring variants that have not run on hardware yet (wctb.h)"""
    check(_lib.load().wctb_debug_set_gram_variant(int(v)), "debug_set_gram_variant")


def eigh_profile(a: torch.Tensor, scale):
    """debug: phase profile (clock64 sums of thread 0) of one C in (64,128] solve -> dict"""
    buf = torch.zeros(16, device=a.device, dtype=torch.int64)
    check(_lib.load().wctb_debug_eigh_profile(buf.data_ptr()), "debug_eigh_profile")
    try:
        eigh_jacobi(a, scale)
        torch.cuda.synchronize()
    finally:
        check(_lib.load().wctb_debug_eigh_profile(None), "debug_eigh_profile")
    v = buf.tolist()
    names = ("load", "cholesky", "sweeps_phase")
    out = dict(zip(names, v[:3]))
    out["sweeps"], out["k"] = v[8], v[9]
    return out


def dp_rate():
    """debug: (cycles per dependent DFMA, DFMA per clock per SM with 16 resident warps)"""
    buf = torch.zeros(4, device="cuda", dtype=torch.int64)
    check(_lib.load().wctb_debug_dp_rate(buf.data_ptr(), _stream()), "debug_dp_rate")
    torch.cuda.synchronize()
    v = buf.tolist()
    return v[0] / 4096.0, 512 * 4096.0 / v[1]


def wct_matrix(c_evals, c_evecs, c_mean, s_evals, s_evecs, s_mean, tau: float, alpha: float, keep_c: int = 0, keep_s: int = 0):
    """-> (M fp32 [C,C], b fp32 [C], mean_c fp32 [C])  with csF = M (cF - mean_c) + b.
    keep_c / keep_s > 0: use only that many of the largest content / style eigen-directions (util_wct.py:26-27,87-88)."""
    C = c_evals.numel()
    dev = c_evals.device
    m = torch.empty(C, C, device=dev, dtype=torch.float32)
    b = torch.empty(C, device=dev, dtype=torch.float32)
    mc = torch.empty(C, device=dev, dtype=torch.float32)
    work = torch.empty(_lib.load().wctb_workspace_doubles(_lib.WS_WCT_MATRIX, C, 1), device=dev, dtype=torch.float64)
    f64 = torch.float64
    if keep_c > 0 or keep_s > 0:
        check(_lib.load().wctb_wct_matrix_topk(_need(c_evals, f64), _need(c_evecs, f64), _need(c_mean, f64), _need(s_evals, f64),
                                               _need(s_evecs, f64), _need(s_mean, f64), C, float(tau), float(alpha), int(keep_c),
                                               int(keep_s), _need(m), _need(b), _need(mc), _need(work, f64), _stream()),
              "wct_matrix_topk")
        _count("wct_matrix")
        return m, b, mc
    check(_lib.load().wctb_wct_matrix(_need(c_evals, f64), _need(c_evecs, f64), _need(c_mean, f64), _need(s_evals, f64),
                                      _need(s_evecs, f64), _need(s_mean, f64), C, float(tau), float(alpha), _need(m),
                                      _need(b), _need(mc), _need(work, f64), _stream()), "wct_matrix")
    _count("wct_matrix")
    return m, b, mc


def whiten_ns(gram: torch.Tensor, scale: float, add_identity: bool = False, return_info: bool = False):
    """fp64 centred Gram [C,C] -> whitening matrix W = (scale*gram [+I])^-1/2 (pseudo-inverse on the range), fp64 [C,C],
    by pivoted Cholesky + Newton-Schulz on a cooperative grid (wctb_whiten_ns; C <= 128).  Opt-in path, see wctb.h."""
    C = gram.shape[-1]
    g = gram.reshape(C, C)
    w = torch.empty(C, C, device=g.device, dtype=torch.float64)
    work = torch.empty(_lib.load().wctb_workspace_doubles(_lib.WS_WHITEN_NS, C, 1), device=g.device, dtype=torch.float64)
    info = torch.zeros(4, device=g.device, dtype=torch.int32)
    check(_lib.load().wctb_whiten_ns(_need(g, torch.float64), float(scale), int(add_identity), C, _need(w, torch.float64),
                                     _need(work, torch.float64), _need(info, torch.int32), _stream()), "whiten_ns")
    _count("whiten_ns")
    return (w, info) if return_info else w


def wct_matrix_w(w_whiten, c_mean, s_evals, s_evecs, s_mean, tau: float, alpha: float):
    """like wct_matrix, with the content side given as a ready whitening matrix (whiten_ns)"""
    C = s_evals.numel()
    dev = s_evals.device
    m = torch.empty(C, C, device=dev, dtype=torch.float32)
    b = torch.empty(C, device=dev, dtype=torch.float32)
    mc = torch.empty(C, device=dev, dtype=torch.float32)
    work = torch.empty(_lib.load().wctb_workspace_doubles(_lib.WS_WCT_MATRIX, C, 1), device=dev, dtype=torch.float64)
    f64 = torch.float64
    check(_lib.load().wctb_wct_matrix_w(_need(w_whiten, f64), _need(c_mean, f64), _need(s_evals, f64), _need(s_evecs, f64),
                                        _need(s_mean, f64), C, float(tau), float(alpha), _need(m), _need(b), _need(mc),
                                        _need(work, f64), _stream()), "wct_matrix_w")
    _count("wct_matrix_w")
    return m, b, mc


def wct_apply(x: torch.Tensor, m: torch.Tensor, b: torch.Tensor, mean_c: torch.Tensor, round_tf32: bool = False) -> torch.Tensor:
    C4, H, W, _ = x.shape
    y = torch.empty_like(x)
    check(_lib.load().wctb_wct_apply(_need(x), _need(m), _need(b), _need(mean_c), _need(y), C4 * 4, H * W, int(round_tf32),
                                     _stream()), "wct_apply")
    _count("wct_apply")
    return y


def fold_wct_into_conv(w_oihw, bias, m, b, mean_c):
    cout, cin = w_oihw.shape[:2]
    w_out = torch.empty_like(w_oihw)
    b_out = torch.empty_like(bias)
    check(_lib.load().wctb_fold_wct_into_conv(_need(w_oihw), _need(bias), _need(m), _need(b), _need(mean_c), _need(w_out),
                                              _need(b_out), cin, cout, _stream()), "fold_wct_into_conv")
    _count("fold")
    return w_out, b_out


# ------------------------------------------------------------------------------------------------ h2 engine
# fp32-accurate tensor-core convolutions: activations travel as H8 = [ceil(C/8), 2 (hi, lo), H, W, 8] fp16 tensors
def h2_supported(cin: int, cout: int) -> bool:
    return bool(_lib.load().wctb_h2_supported(cin, cout))


def _need_h8(t: torch.Tensor):
    if t.dtype != torch.float16 or t.dim() != 5 or t.shape[1] != 2 or t.shape[4] != 8:
        raise _lib.WctbError("expected an H8 activation [C/8,2,H,W,8] fp16, got %s %s" % (t.dtype, tuple(t.shape)))
    return _need(t, torch.float16)


def pack_weights_h2(w_oihw: torch.Tensor):
    """OIHW fp32 -> (packed fp16 halves, wscale fp32[4] device buffer); the power-of-two weight scale is chosen on the device"""
    cout, cin = w_oihw.shape[:2]
    w = w_oihw.detach().contiguous().float()
    n = _lib.load().wctb_h2_packed_halves(cin, cout)
    if n <= 0:
        raise _lib.WctbError("h2 engine does not support a %d -> %d layer" % (cin, cout))
    dst = torch.empty(n, device=w.device, dtype=torch.float16)
    wscale = torch.empty(4, device=w.device, dtype=torch.float32)
    check(_lib.load().wctb_pack_weights_h2(_need(w), _need(dst, torch.float16), _need(wscale), cin, cout, _stream()), "pack_weights_h2")
    _count("pack_h2")
    return dst, wscale


def conv3x3_h2(x_h8: torch.Tensor, w: torch.Tensor, wscale: torch.Tensor, b: torch.Tensor, cin: int, cout: int, epilogue: int,
               out_h8: bool = True, out_p4: bool = False):
    """H8 -> (H8 | None, fp32 P4 | None); EPI_NCHW3 -> (None, image [1,3,H,W])"""
    _, _, H, W, _ = x_h8.shape
    if epilogue == EPI_POOL2:
        Ho, Wo = H // 2, W // 2
    elif epilogue == EPI_UP2:
        Ho, Wo = 2 * H, 2 * W
    else:
        Ho, Wo = H, W
    dev = x_h8.device
    y8 = y4 = None
    if epilogue == EPI_NCHW3:
        y4 = torch.empty(1, 3, H, W, device=dev, dtype=torch.float32)
    else:
        if out_h8:
            y8 = torch.empty((cout + 7) // 8, 2, Ho, Wo, 8, device=dev, dtype=torch.float16)
        if out_p4:
            y4 = torch.empty(cout // 4, Ho, Wo, 4, device=dev, dtype=torch.float32)
    check(_lib.load().wctb_conv3x3_h2(_need_h8(x_h8), _need(w, torch.float16), _need(b), _need(wscale),
                                      None if y8 is None else _need(y8, torch.float16), None if y4 is None else _need(y4),
                                      H, W, cin, cout, epilogue, _stream()), "conv3x3_h2")
    _count("conv_h2")
    return y8, y4


def conv3x3_first_h2(x_nchw: torch.Tensor, w: torch.Tensor, b: torch.Tensor, cout: int, out_h8: bool = True, out_p4: bool = False):
    """x: [1,3,H,W] or [3,H,W]; w packed [9][3][cout] fp32 -> (H8 | None, P4 | None)"""
    if x_nchw.dim() == 4:
        x_nchw = x_nchw.squeeze(0)
    _, H, W = x_nchw.shape
    dev = x_nchw.device
    y8 = torch.empty(cout // 8, 2, H, W, 8, device=dev, dtype=torch.float16) if out_h8 else None
    y4 = torch.empty(cout // 4, H, W, 4, device=dev, dtype=torch.float32) if out_p4 else None
    check(_lib.load().wctb_conv3x3_first_h2(_need(x_nchw), _need(w), _need(b), None if y8 is None else _need(y8, torch.float16),
                                            None if y4 is None else _need(y4), H, W, cout, _stream()), "conv3x3_first_h2")
    _count("conv_first_h2")
    return y8, y4


def nchw_to_h8(x: torch.Tensor) -> torch.Tensor:
    if x.dim() == 4:
        x = x.squeeze(0)
    C, H, W = x.shape
    y = torch.empty((C + 7) // 8, 2, H, W, 8, device=x.device, dtype=torch.float16)
    check(_lib.load().wctb_nchw_to_h8(_need(x), _need(y, torch.float16), C, H, W, _stream()), "nchw_to_h8")
    _count("to_h8")
    return y


def h8_to_nchw(x: torch.Tensor, C: int = None) -> torch.Tensor:
    C8, _, H, W, _ = x.shape
    C = C8 * 8 if C is None else C
    y = torch.empty(1, C, H, W, device=x.device, dtype=torch.float32)
    check(_lib.load().wctb_h8_to_nchw(_need_h8(x), _need(y), C, H, W, _stream()), "h8_to_nchw")
    _count("from_h8")
    return y


def p4_to_h8(x: torch.Tensor) -> torch.Tensor:
    C4, H, W, _ = x.shape
    y = torch.empty((C4 + 1) // 2, 2, H, W, 8, device=x.device, dtype=torch.float16)
    check(_lib.load().wctb_p4_to_h8(_need(x), _need(y, torch.float16), C4 * 4, H, W, _stream()), "p4_to_h8")
    _count("to_h8")
    return y


# ---- fused head / tail of the h2 engine: dx-stacked weight tiles, packed on the host (static weights, once at load)
def h2_host_scale(w: torch.Tensor) -> float:
    """power of two s with max|w| * s in [512, 1024) (same rule as the device-side packer)"""
    import math
    m = float(w.detach().abs().max())
    if not (m > 0.0) or not math.isfinite(m):
        return 1.0
    return 2.0 ** (10 - math.frexp(m)[1])


def _hi_lo(w: torch.Tensor, s: float):
    ws = w.detach().float() * s
    hi = ws.half()
    return hi, (ws - hi.float()).half()


def pack_head_h2_w11(w_oihw: torch.Tensor):
    """conv11 [16,3,3,3] (conv0 folded) -> ([2 mma][2 chunks][96][8] fp16, 1/s).  MMA m, K chunk c carries filter row
    dy = 2m + c (dy = 3 is all zero); a K chunk is one pixel [R G B 0 | r g b 0] (hi | lo); row dx*16+co holds
    [w_hi(RGB) 0 w_hi(RGB) 0] (hi*w_hi + lo*w_hi), row 48+dx*16+co holds [w_lo(RGB) 0 0 0 0 0] (hi*w_lo)."""
    s = h2_host_scale(w_oihw)
    hi, lo = _hi_lo(w_oihw, s)
    t = torch.zeros(2, 2, 96, 8, device=w_oihw.device, dtype=torch.float16)
    for dy in range(3):
        m, c = divmod(dy, 2)
        for dx in range(3):
            r = slice(dx * 16, dx * 16 + 16)
            t[m, c, r, 0:3] = hi[:, :, dy, dx]
            t[m, c, r, 4:7] = hi[:, :, dy, dx]
            t[m, c, 48 + dx * 16:48 + dx * 16 + 16, 0:3] = lo[:, :, dy, dx]
    return t.contiguous(), 1.0 / s


def pack_dx_h2(w_oihw: torch.Tensor):
    """[Cout<=16,16,3,3] -> ([3 dy][2 chunks][96][8] fp16, 1/s): row dx*16+co = w_hi[co, c*8+e, dy, dx], row 48+dx*16+co = w_lo"""
    cout = w_oihw.shape[0]
    assert w_oihw.shape[1] == 16 and cout <= 16
    s = h2_host_scale(w_oihw)
    hi, lo = _hi_lo(w_oihw, s)
    t = torch.zeros(3, 2, 96, 8, device=w_oihw.device, dtype=torch.float16)
    for dy in range(3):
        for c in range(2):
            for dx in range(3):
                t[dy, c, dx * 16:dx * 16 + cout, :] = hi[:, c * 8:c * 8 + 8, dy, dx]
                t[dy, c, 48 + dx * 16:48 + dx * 16 + cout, :] = lo[:, c * 8:c * 8 + 8, dy, dx]
    return t.contiguous(), 1.0 / s


def conv_head_h2(x_nchw: torch.Tensor, w11p, inv_s11: float, b11, w12p, inv_s12: float, b12) -> torch.Tensor:
    """fused conv11+ReLU+conv12+ReLU+pool of the 16x encoders: image [1,3,H,W] -> H8 [2,2,H/2,W/2,8]"""
    if x_nchw.dim() == 4:
        x_nchw = x_nchw.squeeze(0)
    _, H, W = x_nchw.shape
    y = torch.empty(2, 2, H // 2, W // 2, 8, device=x_nchw.device, dtype=torch.float16)
    check(_lib.load().wctb_conv_head_h2(_need(x_nchw), _need(w11p, torch.float16), _need(b11), float(inv_s11),
                                        _need(w12p, torch.float16), _need(b12), float(inv_s12), _need(y, torch.float16),
                                        H, W, _stream()), "conv_head_h2")
    _count("conv_head_h2")
    return y


def conv_tail_h2(x_h8: torch.Tensor, w12p, inv_s12: float, b12, w11p, inv_s11: float, b11, upsample_input: bool,
                 shard: dict = None) -> torch.Tensor:
    """fused [nearest x2 +] conv12(16->16)+ReLU+conv11(16->3)+ReLU of the 16x decoders: H8 [2,2,h,w,8] -> image [1,3,H,W].
    shard (multi-GPU): {"out": next-stage extended strip [1,3,H,We] (written in place and returned), "out_x0", "own_x0", "own_w",
    "halo", "peer_l": (ptr, pitch, x0) | None, "peer_r": ...}: only the own columns are kept; the seam-side `halo` columns also go
    straight into the neighbours' buffers over NVLink (wctb_conv_tail_h2_sharded)."""
    _need_h8(x_h8)
    _, _, h, w, _ = x_h8.shape
    H, W = (2 * h, 2 * w) if upsample_input else (h, w)
    lib = _lib.load()
    if shard is None:
        y = torch.empty(1, 3, H, W, device=x_h8.device, dtype=torch.float32)
        check(lib.wctb_conv_tail_h2(_need(x_h8, torch.float16), _need(w12p, torch.float16), _need(b12), float(inv_s12),
                                    _need(w11p, torch.float16), _need(b11), float(inv_s11), _need(y), H, W,
                                    int(upsample_input), _stream()), "conv_tail_h2")
        _count("conv_tail_h2")
        return y
    out = shard["out"]
    if out.shape[-2] != H or out.dim() != 4 or out.shape[1] != 3:
        raise _lib.WctbError("sharded tail: output strip %s does not match the image height %d" % (tuple(out.shape), H))
    ts = _lib.TailShard()
    ts.out, ts.out_pitch, ts.out_x0 = _need(out), out.shape[-1], int(shard["out_x0"])
    ts.own_x0, ts.own_w, ts.halo = int(shard["own_x0"]), int(shard["own_w"]), int(shard["halo"])
    pl, pr = shard.get("peer_l"), shard.get("peer_r")
    ts.peer_l, ts.peer_l_pitch, ts.peer_l_x0 = (pl[0], pl[1], pl[2]) if pl else (None, 0, 0)
    ts.peer_r, ts.peer_r_pitch, ts.peer_r_x0 = (pr[0], pr[1], pr[2]) if pr else (None, 0, 0)
    import ctypes
    check(lib.wctb_conv_tail_h2_sharded(_need(x_h8, torch.float16), _need(w12p, torch.float16), _need(b12), float(inv_s12),
                                        _need(w11p, torch.float16), _need(b11), float(inv_s11), H, W, int(upsample_input),
                                        ctypes.byref(ts), _stream()), "conv_tail_h2_sharded")
    _count("conv_tail_h2")
    return out
