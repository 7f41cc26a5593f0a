"""ctypes binding of libwctb.so (the C ABI declared in include/wctb.h).

The product path has no CPU fallback: if the library is missing or fails to load, importing
any op raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwctb.so")

_p = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_d = ctypes.c_double

class TailShard(ctypes.Structure):
    """wctb_tail_shard (include/wctb.h): output placement of the strip-sharded fused tail, incl. the neighbours' peer pointers"""
    _fields_ = [("out", ctypes.c_void_p), ("out_pitch", ctypes.c_int), ("out_x0", ctypes.c_int),
                ("own_x0", ctypes.c_int), ("own_w", ctypes.c_int), ("halo", ctypes.c_int),
                ("peer_l", ctypes.c_void_p), ("peer_l_pitch", ctypes.c_int), ("peer_l_x0", ctypes.c_int),
                ("peer_r", ctypes.c_void_p), ("peer_r_pitch", ctypes.c_int), ("peer_r_x0", ctypes.c_int)]


# name -> argtypes ; every function returns int except where noted
SIGNATURES = {
    "wctb_abi_version": [],
    "wctb_last_cuda_error": [],
    "wctb_nchw_to_p4": [_p, _p, _i, _i, _i, _i, _p],
    "wctb_p4_to_nchw": [_p, _p, _i, _i, _i, _p],
    "wctb_pack_weights_fp32": [_p, _p, _i, _i, _p],
    "wctb_pack_weights_tf32": [_p, _p, _i, _i, _p],
    "wctb_tf32_kgroup": [_i, _i],
    "wctb_tf32_supported": [_i, _i],
    "wctb_conv3x3_first": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "wctb_conv3x3_p4": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "wctb_conv3x3_last": [_p, _p, _p, _p, _i, _i, _i, _p],
    "wctb_h2_supported": [_i, _i],
    "wctb_pack_weights_h2": [_p, _p, _p, _i, _i, _p],
    "wctb_conv3x3_h2": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "wctb_conv3x3_first_h2": [_p, _p, _p, _p, _p, _i, _i, _i, _p],
    "wctb_conv_head_h2": [_p, _p, _p, ctypes.c_float, _p, _p, ctypes.c_float, _p, _i, _i, _p],
    "wctb_conv_tail_h2": [_p, _p, _p, ctypes.c_float, _p, _p, ctypes.c_float, _p, _i, _i, _i, _p],
    "wctb_conv_tail_h2_sharded": [_p, _p, _p, ctypes.c_float, _p, _p, ctypes.c_float, _i, _i, _i, ctypes.POINTER(TailShard), _p],
    "wctb_nchw_to_h8": [_p, _p, _i, _i, _i, _p],
    "wctb_h8_to_nchw": [_p, _p, _i, _i, _i, _p],
    "wctb_p4_to_h8": [_p, _p, _i, _i, _i, _p],
    "wctb_conv_head_supported": [_i, _i],
    "wctb_conv_head_tc": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "wctb_conv_tail_supported": [_i, _i],
    "wctb_conv_tail": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "wctb_conv_head": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    "wctb_channel_sum": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p],
    "wctb_centered_gram": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "wctb_centered_gram_fast": [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p],
    "wctb_eigh_jacobi": [_p, _i, _i, ctypes.POINTER(ctypes.c_double), _i, _p, _p, _p, _p, _p],
    "wctb_eigh_jacobi_tol": [_p, _i, _i, ctypes.POINTER(ctypes.c_double), _i, _d, _p, _p, _p, _p, _p],
    "wctb_wct_matrix": [_p, _p, _p, _p, _p, _p, _i, _d, _d, _p, _p, _p, _p, _p],
    "wctb_wct_matrix_topk": [_p, _p, _p, _p, _p, _p, _i, _d, _d, _i, _i, _p, _p, _p, _p, _p],
    "wctb_whiten_ns": [_p, _d, _i, _i, _p, _p, _p, _p],
    "wctb_wct_matrix_w": [_p, _p, _p, _p, _p, _i, _d, _d, _p, _p, _p, _p, _p],
    "wctb_wct_apply": [_p, _p, _p, _p, _p, _i, _ll, _i, _p],
    "wctb_fold_wct_into_conv": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    "wctb_halo_pack": [_p, _p, _i, _i, _i, _i, _i, _p],
    "wctb_halo_unpack": [_p, _p, _i, _i, _i, _i, _i, _p],
    "wctb_u8hwc_to_nchw": [_p, _p, _i, _i, _p],
    "wctb_nchw_to_u8hwc": [_p, _p, _i, _i, _p],
    "wctb_resize_ksize": [_i, _i],
    "wctb_resize_coeffs_host": [_i, _i, _p, _p],
    "wctb_resize_u8_pass": [_p, _p, _i, _i, _i, _i, _p, _p, _i, _p],
    "wctb_debug_set_eigh_variant": [_i],
    "wctb_debug_set_gram_variant": [_i],
    "wctb_debug_set_first_variant": [_i],
    "wctb_debug_eigh_profile": [_p],
    "wctb_debug_dp_rate": [_p, _p],
    "wctb_debug_set_trace": [_p],
    "wctb_debug_mma_rate": [_p, _i, _i, _i, _i, _i, _p],
    "wctb_selftest_umma": [_p, _p, _p, _i, _i, _p],
    "wctb_debug_set_h2_resident": [_i],
    "wctb_debug_mma_rate_f16": [_p, _i, _i, _i, _i, _p],
    "wctb_debug_mma_rate_f16_off": [_p, _i, _i, _i, _i, _i, _p],
    "wctb_debug_ldtm_rate": [_p, _i, _i, _i, _i, _p],
}

WCTB_OK = 0
EPI_NONE, EPI_POOL2, EPI_UP2, EPI_NCHW3 = 0, 1, 2, 3
ENGINE_FP32, ENGINE_TF32, ENGINE_H2 = 0, 1, 2
WS_EIGH, WS_WCT_MATRIX, WS_WHITEN_NS = 0, 1, 2


class WctbError(RuntimeError):
    pass


_lib = None


def load():
    """Load libwctb.so (once).  Raises WctbError if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WctbError("libwctb.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise WctbError("failed to load %s: %s" % (LIB_PATH, e))
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    lib.wctb_error_string.argtypes = [_i]
    lib.wctb_error_string.restype = ctypes.c_char_p
    lib.wctb_workspace_doubles.argtypes = [_i, _i, _i]
    lib.wctb_workspace_doubles.restype = _ll
    lib.wctb_h2_packed_halves.argtypes = [_i, _i]
    lib.wctb_h2_packed_halves.restype = _ll
    if lib.wctb_abi_version() != 1:
        raise WctbError("libwctb ABI version mismatch")
    if os.environ.get("WCTB_H2_RESIDENT"):           # A/B: 0 = stream the weights with every pipeline stage (round-2a kernels)
        lib.wctb_debug_set_h2_resident(int(os.environ["WCTB_H2_RESIDENT"]))
    if os.environ.get("WCTB_GRAM_VARIANT"):          # A/B switch for tools / bench runs (see wctb.h, debug section)
        lib.wctb_debug_set_gram_variant(int(os.environ["WCTB_GRAM_VARIANT"]))
    if os.environ.get("WCTB_FIRST_VARIANT"):
        lib.wctb_debug_set_first_variant(int(os.environ["WCTB_FIRST_VARIANT"]))
    _lib = lib
    return lib


# ---- libwctb_io.so: nvJPEG behind include/wctb_io.h (stateful codec; separate library) ----------------------
IO_LIB_PATH = os.path.join(_HERE, "csrc", "libwctb_io.so")
_sz = ctypes.c_size_t
_psz = ctypes.POINTER(ctypes.c_size_t)
_pi = ctypes.POINTER(ctypes.c_int)
IO_SIGNATURES = {
    "wctb_io_abi_version": [],
    "wctb_io_last_status": [],
    "wctb_io_create": [ctypes.POINTER(_p)],
    "wctb_io_create_ex": [_i, ctypes.c_uint, ctypes.POINTER(_p)],
    "wctb_io_jpeg_info": [_p, _p, _sz, _pi, _pi, _pi, _pi],
    "wctb_io_jpeg_decode": [_p, _p, _sz, _p, _i, _i, _p],
    "wctb_io_jpeg_encode": [_p, _p, _i, _i, _i, _i, _p, _psz],
    "wctb_io_jpeg_retrieve": [_p, _p, _sz, _psz, _p],
}
_io_lib = None


def load_io():
    """Load libwctb_io.so (once).  Raises WctbError if it is missing or nvJPEG cannot be resolved."""
    global _io_lib
    if _io_lib is not None:
        return _io_lib
    if not os.path.exists(IO_LIB_PATH):
        raise WctbError("libwctb_io.so not found at %s -- run the build (csrc/build.py)" % IO_LIB_PATH)
    try:
        lib = ctypes.CDLL(IO_LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise WctbError("failed to load %s: %s" % (IO_LIB_PATH, e))
    for name, args in IO_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    lib.wctb_io_destroy.argtypes = [_p]
    lib.wctb_io_destroy.restype = None
    lib.wctb_io_error_string.argtypes = [_i]
    lib.wctb_io_error_string.restype = ctypes.c_char_p
    if lib.wctb_io_abi_version() != 1:
        raise WctbError("libwctb_io ABI version mismatch")
    _io_lib = lib
    return lib


def check_io(code: int, what: str):
    if code != 0:
        lib = load_io()
        raise WctbIoError(code, "%s failed: %s (status %d)" % (what, lib.wctb_io_error_string(code).decode(), lib.wctb_io_last_status()))


class WctbIoError(WctbError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def check(code: int, what: str):
    if code != WCTB_OK:
        lib = load()
        msg = lib.wctb_error_string(code).decode()
        if code == -4:
            msg += " (cudaError %d)" % lib.wctb_last_cuda_error()
        raise WctbError("%s failed: %s" % (what, msg))
