"""CPU oracle for the image I/O around the WCT hot path  --  TEST INFRASTRUCTURE ONLY.

Restates, in numpy integer / IEEE arithmetic, what the reference does to an image before and after the hot path:

  PytorchWCT/data_loader.py:52-55   transforms.Resize(size)(PIL image)   -> shorter side = size, bilinear, antialiased
  PytorchWCT/data_loader.py:56-57   transforms.ToTensor()                 -> u8 HWC -> fp32 CHW, x / 255
  PytorchWCT/WCT.py:128             vutils.save_image(img)                -> fp32 CHW -> u8 HWC, trunc(clamp(x*255 + 0.5, 0, 255))

The arithmetic lives in third-party dependencies that are not under /root/reference: Pillow (requirements.txt pins
`Pillow==6.2.2`; this image has 12.2.0 -- the 8-bit resampler `src/libImaging/Resample.c` is unchanged between them:
double-precision triangle-filter coefficients normalised per output pixel, converted to 22-bit fixed point
(`PRECISION_BITS = 32 - 8 - 2`), horizontal pass then vertical pass with an 8-bit intermediate image, each pass
rounding with `(acc + 2^21) >> 22` and clamping to [0,255]) and torchvision (`transforms.functional.resize` size rule,
`ToTensor`, `utils.save_image`).  Parity pinning: the reference has no fixtures for this either; both libraries are
importable wherever the tests run (same image on the GPU box), so `tests/test_image_io_host.py` pins this file
bit-exactly against live PIL / torchvision calls, and the CUDA kernels are pinned bit-exactly against this file.

Only tests, smoke() and bench.py's CPU legs may import this module; the product never does.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2          # Resample.c: 8 bits of pixel, 2 bits of head-room for the accumulation


def resized_output_size(h: int, w: int, size: int):
    """torchvision.transforms.functional._compute_resized_output_size for an int `size`:
    shorter side -> size, longer side -> int(size * long / short)   (data_loader.py:52-55 via transforms.Resize)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    new_w, new_h = (new_short, new_long) if w <= h else (new_long, new_short)
    return new_h, new_w


def texture_output_size(h: int, w: int, size: int):
    """data_loader.py:64-72 (texture synthesis): LONGER side -> size (note: opposite of transforms.Resize)."""
    if w > h:
        return int(h * size / w), size
    return size, int(w * size / h)


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0) and the full box.
    -> (ksize, bounds int32 [out,2] = (xmin, count), coeffs int32 [out,ksize] in 22-bit fixed point)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            wv = 1.0 - a if a < 1.0 else 0.0
            k[x] = wv
            ww += wv
        if ww != 0.0:
            k[:xmax] = k[:xmax] / ww
        bounds[xx] = (xmin, xmax)
        # normalize_coeffs_8bpc: round half away from zero (bilinear weights are >= 0)
        coeffs[xx] = np.where(k < 0, -0.5 + k * (1 << PRECISION_BITS), 0.5 + k * (1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)
    return ksize, bounds, coeffs


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one 8-bit resampling pass along `axis` of an HWC u8 image"""
    in_size = img.shape[axis]
    _, bounds, coeffs = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.tensordot(coeffs[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_u8(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), BILINEAR)` on an RGB image: horizontal pass first (only if the width changes),
    then the vertical pass (only if the height changes); identical size returns a copy (Image.py resize)."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    h, w, _ = img_hwc.shape
    out = img_hwc
    if out_w != w:
        out = _pass(out, out_w, 1)
    if out_h != h:
        out = _pass(out, out_h, 0)
    return out.copy() if out is img_hwc else out


def to_tensor(img_hwc: np.ndarray) -> np.ndarray:
    """transforms.ToTensor on an RGB PIL image: u8 HWC -> fp32 CHW, correctly rounded fp32 division by 255."""
    return (img_hwc.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def save_image_quantize(img_chw: np.ndarray) -> np.ndarray:
    """torchvision.utils.save_image for one image: `grid.mul(255).add_(0.5).clamp_(0, 255).permute(1,2,0).to(uint8)`
    -- two separately rounded fp32 operations (no FMA), clamp, truncation toward zero.  fp32 CHW -> u8 HWC."""
    x = img_chw.astype(np.float32)
    x = (x * np.float32(255.0)).astype(np.float32)
    x = (x + np.float32(0.5)).astype(np.float32)
    x = np.clip(x, np.float32(0.0), np.float32(255.0))
    return x.astype(np.uint8).transpose(1, 2, 0).copy()
