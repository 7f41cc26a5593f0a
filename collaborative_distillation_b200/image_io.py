"""Image I/O around the stylization path, on the GPU.

Mirrors what the reference does on the host before and after the hot path:

  PytorchWCT/data_loader.py:17-18  Image.open(path).convert('RGB')   -> `JpegCodec.decode` (nvJPEG; PNG / CMYK: PIL decode)
  PytorchWCT/data_loader.py:52-55  transforms.Resize(size)            -> `resize_u8` (bit-exact with PIL's 8-bit resampler)
  PytorchWCT/data_loader.py:56-57  transforms.ToTensor()              -> `to_tensor`  (u8 / 255, bit-exact)
  PytorchWCT/WCT.py:128            vutils.save_image(img, path)       -> `quantize` (bit-exact) + `JpegCodec.encode`

Every pixel operation runs in libwctb.so kernels on CUDA tensors (no CPU fallback for them); only container
parsing that nvJPEG does not cover (PNG, CMYK JPEG) is delegated to PIL, after which the 8-bit image is uploaded
and follows the same device path.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import WctbError, WctbIoError, check, check_io
from .ops import _need, _stream

_counts = {"resize_pass": 0, "to_tensor": 0, "quantize": 0, "jpeg_decode": 0, "jpeg_encode": 0}


def launch_counts() -> dict:
    return dict(_counts)


# ------------------------------------------------------------------------------------------ size rules (host, exact)
def resized_output_size(h: int, w: int, size: int):
    """transforms.Resize(int): shorter side -> size, longer side -> int(size * long / short)  (data_loader.py:52-55)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = int(size), int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)          # (new_h, new_w)


def texture_output_size(h: int, w: int, size: int):
    """data_loader.py:64-72 (--synthesis): LONGER side -> size."""
    if w > h:
        return int(h * size / w), int(size)
    return int(size), int(w * size / h)


# ------------------------------------------------------------------------------------------ resize
_coeff_cache = {}


def resize_coeffs_host(in_size: int, out_size: int):
    """(ksize, bounds int32 [out,2], coeffs int32 [out,ksize]) from wctb_resize_coeffs_host -- host only, no GPU."""
    lib = _lib.load()
    ksize = lib.wctb_resize_ksize(int(in_size), int(out_size))
    check(ksize if ksize < 0 else 0, "resize_ksize")
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    check(lib.wctb_resize_coeffs_host(int(in_size), int(out_size), bounds.ctypes.data, coeffs.ctypes.data), "resize_coeffs_host")
    return ksize, bounds, coeffs


def _coeffs_device(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    ent = _coeff_cache.get(key)
    if ent is None:
        ksize, bounds, coeffs = resize_coeffs_host(in_size, out_size)
        ent = (ksize, torch.from_numpy(bounds).to(device), torch.from_numpy(coeffs).to(device))
        if len(_coeff_cache) > 64:
            _coeff_cache.clear()
        _coeff_cache[key] = ent
    return ent


def _need_u8(t: torch.Tensor):
    if not t.is_cuda:
        raise WctbError("image kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.uint8 or not t.is_contiguous() or t.dim() != 3 or t.shape[2] != 3:
        raise WctbError("expected a contiguous uint8 [H,W,3] tensor, got %s %s" % (t.dtype, tuple(t.shape)))
    return t.data_ptr()


def resize_u8(img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """u8 [H,W,3] CUDA -> u8 [out_h,out_w,3]: PIL `resize((out_w,out_h), BILINEAR)` (horizontal pass, then vertical)."""
    _need_u8(img)
    lib = _lib.load()
    H, W, _ = img.shape
    cur = img
    for axis, n_in, n_out in ((1, W, int(out_w)), (0, H, int(out_h))):
        if n_in == n_out:
            continue
        ksize, bounds, coeffs = _coeffs_device(n_in, n_out, img.device)
        h, w, _ = cur.shape
        dst = torch.empty((h, n_out, 3) if axis == 1 else (n_out, w, 3), dtype=torch.uint8, device=img.device)
        check(lib.wctb_resize_u8_pass(_need_u8(cur), _need_u8(dst), h, w, n_out, axis, _need(bounds, torch.int32),
                                      _need(coeffs, torch.int32), ksize, _stream()), "resize_u8_pass")
        _counts["resize_pass"] += 1
        cur = dst
    return cur.clone() if cur is img else cur


def to_tensor(img: torch.Tensor) -> torch.Tensor:
    """u8 [H,W,3] CUDA -> fp32 [1,3,H,W] in [0,1] (ToTensor)."""
    H, W, _ = img.shape
    out = torch.empty(1, 3, H, W, dtype=torch.float32, device=img.device)
    check(_lib.load().wctb_u8hwc_to_nchw(_need_u8(img), _need(out), H, W, _stream()), "u8hwc_to_nchw")
    _counts["to_tensor"] += 1
    return out


def quantize(img: torch.Tensor) -> torch.Tensor:
    """fp32 [1,3,H,W] (or [3,H,W]) CUDA -> u8 [H,W,3]: the rounding of vutils.save_image (WCT.py:128)."""
    if img.dim() == 4:
        if img.shape[0] != 1:
            raise WctbError("quantize takes one image (the reference saves batch-1 tensors)")
        img = img[0]
    if img.shape[0] != 3:
        raise WctbError("expected 3 channels")
    img = img.contiguous()
    _, H, W = img.shape
    out = torch.empty(H, W, 3, dtype=torch.uint8, device=img.device)
    check(_lib.load().wctb_nchw_to_u8hwc(_need(img), _need_u8(out), H, W, _stream()), "nchw_to_u8hwc")
    _counts["quantize"] += 1
    return out


# ------------------------------------------------------------------------------------------ JPEG codec (nvJPEG)
SUBSAMPLING = {"444": 0, "4:4:4": 0, "422": 1, "4:2:2": 1, "420": 2, "4:2:0": 2}


class JpegCodec:
    """One nvJPEG handle + decoder / encoder state (include/wctb_io.h).  Not thread-safe: one per host thread."""

    BACKENDS = {"default": 0, "hybrid": 1, "gpu_hybrid": 2}

    def __init__(self, backend: str = "default", interp_upsampling: bool = False):
        """backend / interp_upsampling other than the defaults go through wctb_io_create_ex (not yet run on hardware)."""
        self._lib = _lib.load_io()
        h = ctypes.c_void_p()
        if backend == "default" and not interp_upsampling:
            check_io(self._lib.wctb_io_create(ctypes.byref(h)), "io_create")
        else:
            check_io(self._lib.wctb_io_create_ex(self.BACKENDS[backend], 1 if interp_upsampling else 0, ctypes.byref(h)), "io_create_ex")
        self._h = h

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.wctb_io_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass

    def info(self, data: bytes):
        """-> (height, width, components, nvjpeg subsampling enum)"""
        w, h, c, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
        check_io(self._lib.wctb_io_jpeg_info(self._h, buf, len(data), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c),
                                             ctypes.byref(s)), "jpeg_info")
        return h.value, w.value, c.value, s.value

    def decode(self, data: bytes, device="cuda") -> torch.Tensor:
        """JPEG bytes -> u8 [H,W,3] RGB on the device."""
        h, w, _, _ = self.info(data)
        out = torch.empty(h, w, 3, dtype=torch.uint8, device=device)
        buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
        check_io(self._lib.wctb_io_jpeg_decode(self._h, buf, len(data), _need_u8(out), w, h, _stream()), "jpeg_decode")
        torch.cuda.current_stream().synchronize()     # `buf` (host) must outlive nvJPEG's asynchronous copies
        _counts["jpeg_decode"] += 1
        return out

    def encode(self, img: torch.Tensor, quality: int = 75, subsampling: str = "420") -> bytes:
        """u8 [H,W,3] RGB on the device -> baseline JPEG bytes (PIL's save defaults: quality 75, 4:2:0)."""
        H, W, _ = img.shape
        n = ctypes.c_size_t()
        st = _stream()
        check_io(self._lib.wctb_io_jpeg_encode(self._h, _need_u8(img), W, H, int(quality), SUBSAMPLING[subsampling], st,
                                               ctypes.byref(n)), "jpeg_encode")
        out = (ctypes.c_ubyte * n.value)()
        m = ctypes.c_size_t()
        check_io(self._lib.wctb_io_jpeg_retrieve(self._h, out, n.value, ctypes.byref(m), st), "jpeg_retrieve")
        _counts["jpeg_encode"] += 1
        return ctypes.string_at(out, m.value)


_default_codec = None


def default_codec() -> JpegCodec:
    global _default_codec
    if _default_codec is None:
        _default_codec = JpegCodec()
    return _default_codec


def _pil_decode_u8(path) -> torch.Tensor:
    from PIL import Image
    return torch.from_numpy(np.asarray(Image.open(path).convert("RGB")).copy())


def decode_file(path: str, codec: JpegCodec = None, device="cuda") -> torch.Tensor:
    """file -> u8 [H,W,3] on the device.  JPEG goes through nvJPEG; other containers (PNG) and JPEG variants nvJPEG
    refuses (CMYK) are parsed by PIL on the host and uploaded as 8-bit RGB."""
    if path.lower().endswith((".jpg", ".jpeg")):
        with open(path, "rb") as f:
            data = f.read()
        try:
            return (codec or default_codec()).decode(data, device)
        except WctbIoError as e:
            if e.code != -2:          # WCTB_IO_E_UNSUPPORTED
                raise
    return _pil_decode_u8(path).to(device)


def load_image(path: str, size: int = 0, codec: JpegCodec = None, device="cuda", longer_side: bool = False) -> torch.Tensor:
    """data_loader.py:46-57 on the device: decode, optional Resize(size), ToTensor -> fp32 [1,3,H,W].
    longer_side=True applies the texture rule of data_loader.py:64-72 instead of transforms.Resize's."""
    img = decode_file(path, codec, device)
    if size:
        H, W, _ = img.shape
        oh, ow = (texture_output_size if longer_side else resized_output_size)(H, W, size)
        img = resize_u8(img, oh, ow)
    return to_tensor(img)


def save_image(img: torch.Tensor, path: str, codec: JpegCodec = None, quality: int = 75, subsampling: str = "420"):
    """vutils.save_image(img, path) for one image (WCT.py:128): quantise on the device; .jpg/.jpeg are encoded by nvJPEG
    with PIL's defaults, other extensions are written by PIL from the downloaded 8-bit image."""
    q = quantize(img)
    if path.lower().endswith((".jpg", ".jpeg")):
        data = (codec or default_codec()).encode(q, quality, subsampling)
        with open(path, "wb") as f:
            f.write(data)
        return
    from PIL import Image
    Image.fromarray(q.cpu().numpy()).save(path)
