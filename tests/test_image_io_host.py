"""CPU-only checks of the image-I/O row: the numpy oracle is pinned bit-exactly against live PIL /
torchvision calls (the third-party libraries the reference's data_loader.py / save_image use), the C host coefficient
routine of libwctb.so is pinned against the oracle, and libwctb_io.so exports what include/wctb_io.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from collaborative_distillation_b200 import _lib, image_io
from oracle import image_io_oracle as IO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RESIZE_CASES = [(37, 53, 20), (64, 48, 100), (101, 67, 33), (200, 300, 64), (33, 33, 33), (90, 160, 89), (17, 400, 16),
                (480, 270, 512), (5, 7, 3), (1000, 30, 10), (216, 384, 108)]


def _noise(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,size", RESIZE_CASES)
def test_oracle_resize_equals_torchvision_resize_on_pil(h, w, size):
    """data_loader.py:52-55: transforms.Resize(size)(PIL image) -- bit-exact, including the output-size rule."""
    from PIL import Image
    import torchvision.transforms as T
    img = _noise(h, w, seed=h * 1000 + w)
    ref = np.asarray(T.Resize(size)(Image.fromarray(img)))
    oh, ow = IO.resized_output_size(h, w, size)
    assert (oh, ow) == ref.shape[:2] == image_io.resized_output_size(h, w, size)
    assert np.array_equal(IO.resize_u8(img, oh, ow), ref)


@pytest.mark.parametrize("h,w,oh,ow", [(40, 50, 13, 77), (40, 50, 80, 20), (123, 77, 123, 30), (123, 77, 60, 77), (300, 200, 7, 9),
                                       (31, 29, 1, 1), (2, 2, 9, 9)])
def test_oracle_resize_equals_pil_for_free_sizes(h, w, oh, ow):
    from PIL import Image
    img = _noise(h, w, seed=7)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(IO.resize_u8(img, oh, ow), ref)


def test_texture_size_rule():
    # data_loader.py:64-72: longer side -> style_size (ints truncated)
    assert IO.texture_output_size(300, 400, 256) == (192, 256) == image_io.texture_output_size(300, 400, 256)
    assert IO.texture_output_size(401, 300, 256) == (256, 191) == image_io.texture_output_size(401, 300, 256)
    assert IO.texture_output_size(300, 300, 100) == (100, 100)


def test_oracle_to_tensor_and_save_image_quantisation():
    from PIL import Image
    import torchvision.transforms as T
    from torchvision.utils import make_grid
    img = _noise(19, 23)
    img[0, :3, 0] = (0, 1, 255)
    assert np.array_equal(T.ToTensor()(Image.fromarray(img)).numpy(), IO.to_tensor(img))
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 37, 41, generator=g) * 1.6 - 0.2           # the decoder output is >= 0 but not clamped above
    x[0, 0, 0, :8] = torch.tensor([0.0, 1.0, 0.5, 254.5 / 255, 0.49999 / 255, 1.5 / 255, -1.0, 2.0])
    # utils.save_image: make_grid (identity for one image) then this exact chain
    ref = make_grid(x).mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy()
    assert np.array_equal(IO.save_image_quantize(x[0].numpy()), ref)
    # ToTensor -> save_image is the identity on 8-bit images
    assert np.array_equal(IO.save_image_quantize(IO.to_tensor(img)), img)


@pytest.mark.parametrize("n_in,n_out", [(53, 28), (28, 53), (100, 100), (3840, 512), (2160, 288), (7, 1), (1, 7), (1000, 333),
                                        (4096, 4095), (10240, 1024), (17, 16), (16, 17)])
def test_c_coefficients_equal_oracle(n_in, n_out):
    """wctb_resize_coeffs_host (what the device pass consumes) == Pillow's precompute_coeffs + normalize_coeffs_8bpc."""
    ksize, bounds, coeffs = image_io.resize_coeffs_host(n_in, n_out)
    k2, b2, c2 = IO.precompute_coeffs(n_in, n_out)
    assert ksize == k2
    assert np.array_equal(bounds, b2)
    assert np.array_equal(coeffs, c2)
    # the taps of every output pixel lie inside the input and sum to 2^22 up to rounding
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all() and (bounds[:, 1] >= 1).all()
    assert np.abs(coeffs.sum(1) - (1 << 22)).max() <= ksize


def test_c_coefficients_reproduce_pil_through_an_integer_pass():
    """the int32 arithmetic the CUDA pass performs, replayed in numpy with the C-computed tables, equals PIL."""
    from PIL import Image
    img = _noise(61, 97, seed=5)
    oh, ow = 23, 40
    cur = img
    for axis, n_out in ((1, ow), (0, oh)):
        ksize, bounds, coeffs = image_io.resize_coeffs_host(cur.shape[axis], n_out)
        src = np.moveaxis(cur, axis, 0).astype(np.int32)
        out = np.empty((n_out,) + src.shape[1:], np.uint8)
        for xx in range(n_out):
            x0, n = bounds[xx]
            acc = np.full(src.shape[1:], 1 << 21, np.int32)
            for i in range(n):
                acc = acc + coeffs[xx, i] * src[x0 + i]
            out[xx] = np.clip(acc >> 22, 0, 255)
        cur = np.moveaxis(out, 0, axis)
    assert np.array_equal(cur, np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    assert lib.wctb_resize_ksize(0, 5) == -1 and lib.wctb_resize_ksize(5, 0) == -1
    assert lib.wctb_resize_coeffs_host(4, 4, None, None) == -1
    assert lib.wctb_u8hwc_to_nchw(None, None, 4, 4, None) == -1
    assert lib.wctb_nchw_to_u8hwc(None, None, 4, 4, None) == -1
    assert lib.wctb_resize_u8_pass(None, None, 4, 4, 2, 1, None, None, 3, None) == -1
    with pytest.raises(_lib.WctbError):
        image_io.to_tensor(torch.zeros(4, 4, 3, dtype=torch.uint8))          # CPU tensor: no fallback
    with pytest.raises(_lib.WctbError):
        image_io.resize_u8(torch.zeros(4, 4, 3, dtype=torch.uint8), 2, 2)


def test_io_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wctb_io.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wctb_io_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) == 10
    lib = ctypes.CDLL(_lib.IO_LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libwctb_io.so does not export %s" % name
    assert declared == set(_lib.IO_SIGNATURES) | {"wctb_io_error_string", "wctb_io_destroy"}
    io = _lib.load_io()
    assert io.wctb_io_abi_version() == 1
    assert io.wctb_io_error_string(-2) == b"JPEG variant not supported by the GPU decoder"
    assert io.wctb_io_jpeg_info(None, None, 0, None, None, None, None) == -1
    assert io.wctb_io_create_ex(0, 0, None) == -1


def test_device_loader_pairs_like_the_reference_dataset(tmp_path):
    """--gpu_io iterates the same content x style pairs, names and size arguments as DataLoader over Dataset
    (reference data_loader.py:32-36,46-59); the decode itself is stubbed here (no GPU)."""
    import importlib.util
    import sys
    from PIL import Image
    cdir, sdir = tmp_path / "content", tmp_path / "style"
    cdir.mkdir(), sdir.mkdir()
    for d, names in ((cdir, ["a.jpg", "b.png", "skip.txt"]), (sdir, ["s1.jpg", "s2.jpeg"])):
        for n in names:
            if n.endswith(".txt"):
                (d / n).write_text("x")
            else:
                Image.fromarray(_noise(8, 9)).save(d / n)
    sys.path.insert(0, os.path.join(ROOT, "PytorchWCT"))
    try:
        spec = importlib.util.spec_from_file_location("wct_cli_io", os.path.join(ROOT, "PytorchWCT", "WCT.py"))
        cli = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cli)
        from data_loader import Dataset
    finally:
        sys.path.pop(0)
    ds = Dataset(str(cdir), str(sdir), str(sdir), 0, 6)
    calls = []

    class FakeIO:
        @staticmethod
        def load_image(path, size=0, longer_side=False):
            calls.append((os.path.basename(path), size, longer_side))
            return torch.zeros(1, 3, 4, 4)

    got = [(tuple(c.shape), name[0]) for c, s, name in cli._DeviceLoader(ds, FakeIO)]
    ref = [ds[i][2] for i in range(len(ds))]
    assert [n for _, n in got] == ref and len(ref) == 4 and all(sh == (1, 3, 4, 4) for sh, _ in got)
    assert calls[0][1:] == (0, False) and calls[1][1:] == (6, False)
    assert cli.parse(["--gpu_io"]).gpu_io and not cli.parse([]).gpu_io
    syn = Dataset(str(cdir), str(sdir), str(sdir), 0, 5, synthesis=True)
    calls.clear()
    items = list(cli._DeviceLoader(syn, FakeIO))
    assert len(items) == 2 and calls[0][1:] == (5, True) and items[0][2][0] in ("s1.jpg", "s2.jpg")
