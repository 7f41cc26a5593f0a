"""GPU parity of the image-I/O row through the C ABI: byte / integer kernels bit-exact against the
oracle (itself pinned against PIL / torchvision in tests/test_image_io_host.py) and against live PIL; nvJPEG decode /
encode against PIL's libjpeg within a stated tolerance."""
import io
import os

import numpy as np
import pytest
import torch

from collaborative_distillation_b200 import image_io
from oracle import image_io_oracle as IO

pytestmark = pytest.mark.gpu      # first green B200 run: profiles/r01_io_hw_check.txt


def _noise(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def _smooth(h, w):
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127.5 + 120 * np.sin(x * 0.021) * np.cos(y * 0.017), 127.5 + 120 * np.sin(x * 0.013 + y * 0.011),
                    255.0 * x / w * y / h], -1)
    return img.astype(np.uint8)


@pytest.mark.parametrize("h,w,size", [(37, 53, 20), (64, 48, 100), (101, 67, 33), (200, 300, 64), (33, 33, 33), (90, 160, 89),
                                      (17, 400, 16), (480, 270, 512), (5, 7, 3), (1000, 30, 10)])
def test_resize_bit_exact_vs_oracle_and_pil(h, w, size):
    from PIL import Image
    import torchvision.transforms as T
    img = _noise(h, w, seed=h + w)
    oh, ow = image_io.resized_output_size(h, w, size)
    got = image_io.resize_u8(torch.from_numpy(img).cuda(), oh, ow).cpu().numpy()
    assert np.array_equal(got, IO.resize_u8(img, oh, ow))
    assert np.array_equal(got, np.asarray(T.Resize(size)(Image.fromarray(img))))


@pytest.mark.parametrize("h,w,oh,ow", [(40, 50, 13, 77), (40, 50, 80, 20), (123, 77, 123, 30), (123, 77, 60, 77), (31, 29, 1, 1),
                                       (2, 2, 9, 9)])
def test_resize_free_sizes_bit_exact(h, w, oh, ow):
    img = _noise(h, w, seed=3)
    got = image_io.resize_u8(torch.from_numpy(img).cuda(), oh, ow).cpu().numpy()
    assert np.array_equal(got, IO.resize_u8(img, oh, ow))


def test_to_tensor_and_quantize_bit_exact():
    img = _noise(67, 129)
    t = image_io.to_tensor(torch.from_numpy(img).cuda())
    assert t.shape == (1, 3, 67, 129) and t.dtype == torch.float32
    assert np.array_equal(t[0].cpu().numpy(), IO.to_tensor(img))
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, 3, 45, 83, generator=g) * 1.6 - 0.2
    x[0, 1, 0, :8] = torch.tensor([0.0, 1.0, 0.5, 254.5 / 255, 0.49999 / 255, 1.5 / 255, -1.0, 2.0])
    q = image_io.quantize(x.cuda()).cpu().numpy()
    assert np.array_equal(q, IO.save_image_quantize(x[0].numpy()))
    # round trip: 8-bit image -> ToTensor -> save_image quantisation is the identity
    assert np.array_equal(image_io.quantize(t).cpu().numpy(), img)


def test_full_size_properties_uhd():
    """BASELINE cfg3 / cfg4 sized images: size-independent properties instead of a slow CPU comparison."""
    h, w = 2160, 3840
    img = torch.randint(0, 256, (h, w, 3), dtype=torch.uint8, device="cuda")
    t = image_io.to_tensor(img)
    assert torch.equal(image_io.quantize(t), img)                      # identity on 8-bit data
    assert torch.equal(image_io.resize_u8(img, h, w), img)             # same size = copy
    const = torch.full((h, w, 3), 117, dtype=torch.uint8, device="cuda")
    r = image_io.resize_u8(const, 1080, 1920)
    assert r.shape == (1080, 1920, 3) and bool((r == 117).all())       # normalised weights reproduce constants
    # rows of a column-constant image stay constant under a width change
    col = torch.arange(h, device="cuda", dtype=torch.int32).remainder(251).to(torch.uint8).view(h, 1, 1).expand(h, w, 3).contiguous()
    r = image_io.resize_u8(col, h, 512)
    assert torch.equal(r, col[:, :512].contiguous())


def test_jpeg_decode_close_to_pil():
    """nvJPEG vs libjpeg (PIL) on the same bitstream: the IDCT may differ by +-1 and nvJPEG does not use libjpeg's
    'fancy' chroma upsampling, so 4:4:4 must agree to <= 4 levels (mean <= 1), 4:2:0 to a mean of <= 1.5 levels (smooth image)."""
    from PIL import Image
    img = _smooth(360, 500)
    codec = image_io.JpegCodec()
    for sub, max_tol, mean_tol in ((0, 4, 1.0), (2, 32, 1.5)):
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", quality=90, subsampling=sub)
        data = buf.getvalue()
        assert codec.info(data)[:3] == (360, 500, 3)
        got = codec.decode(data).cpu().numpy().astype(np.int32)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
        d = np.abs(got - ref)
        assert d.max() <= max_tol and d.mean() <= mean_tol, (sub, d.max(), d.mean())
    # grayscale files expand to RGB like convert('RGB')
    buf = io.BytesIO()
    Image.fromarray(img[..., 0]).save(buf, format="JPEG", quality=90)
    got = codec.decode(buf.getvalue()).cpu().numpy().astype(np.int32)
    ref = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.int32)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 2
    codec.close()


def test_jpeg_encode_is_readable_by_pil_and_close_to_pil_encode():
    from PIL import Image
    img = _smooth(360, 500)
    codec = image_io.JpegCodec()
    data = codec.encode(torch.from_numpy(img).cuda(), quality=75, subsampling="420")
    back = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.float64)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG")            # PIL defaults: quality 75, 4:2:0 (what save_image uses)
    pil_back = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.float64)
    psnr = lambda a: 10 * np.log10(255.0 ** 2 / np.mean((a - img) ** 2))
    assert back.shape == img.shape
    # measured on B200 (profiles/r01_io_hw_check.txt): nvJPEG 42.5 dB vs libjpeg 46.4 dB on this smooth image at the same
    # nominal quality (different chroma downsampling filter and quantisation-table scaling) -- same class, not equal
    assert psnr(back) >= 38.0 and psnr(back) >= psnr(pil_back) - 6.0, (psnr(back), psnr(pil_back))
    assert 0.5 <= len(data) / len(buf.getvalue()) <= 2.0
    codec.close()


def test_load_and_save_image_mirror_the_reference_loader(tmp_path):
    """data_loader.py:46-57 (decode, Resize, ToTensor) and WCT.py:128 (save_image) through the device path."""
    from PIL import Image
    import torchvision.transforms as T
    img = _smooth(300, 420)
    png = os.path.join(tmp_path, "c.png")
    Image.fromarray(img).save(png)
    ref = T.ToTensor()(T.Resize(128)(Image.open(png).convert("RGB")))
    got = image_io.load_image(png, 128)
    assert got.shape == (1, 3, 128, 179) and torch.equal(got[0].cpu(), ref)          # PNG: lossless container -> bit-exact
    jpg = os.path.join(tmp_path, "c.jpg")
    Image.fromarray(img).save(jpg, quality=92)
    got = image_io.load_image(jpg, 0)
    ref = T.ToTensor()(Image.open(jpg).convert("RGB"))
    assert got.shape[1:] == ref.shape and (got[0].cpu() - ref).abs().mean().item() <= 1.5 / 255
    out_png, out_jpg = os.path.join(tmp_path, "o.png"), os.path.join(tmp_path, "o.jpg")
    x = torch.rand(1, 3, 64, 96, device="cuda") * 1.3
    image_io.save_image(x, out_png)
    assert np.array_equal(np.asarray(Image.open(out_png)), IO.save_image_quantize(x[0].cpu().numpy()))
    image_io.save_image(got, out_jpg)
    back = np.asarray(Image.open(out_jpg).convert("RGB")).astype(np.float64)
    assert back.shape == (300, 420, 3) and np.abs(back - img).mean() <= 4.0


@pytest.mark.parametrize("backend,interp", [("gpu_hybrid", False), ("hybrid", False), ("default", True)])
def test_jpeg_decode_other_backends(backend, interp):
    """wctb_io_create_ex: GPU-assisted Huffman backend / interpolated chroma upsampling decode the same picture"""
    from PIL import Image
    img = _smooth(360, 500)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=90, subsampling=2)
    ref = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB")).astype(np.int32)
    codec = image_io.JpegCodec(backend=backend, interp_upsampling=interp)
    got = codec.decode(buf.getvalue()).cpu().numpy().astype(np.int32)
    d = np.abs(got - ref)
    assert d.mean() <= 1.5, (backend, interp, d.max(), d.mean())
    codec.close()
