"""tcgen05 TF32 engine: descriptor self-test and conv parity.  Inputs and weights are pre-rounded to TF32, so every
product is exact in fp32 and the only difference to an fp64 reference is fp32 accumulation order: tight tolerances
that catch any indexing / descriptor / pipeline error."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from collaborative_distillation_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def tf32_rna(t: torch.Tensor) -> torch.Tensor:
    """round-to-nearest, ties away from zero, to 10 mantissa bits (cvt.rna.tf32.f32)"""
    u = t.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    u = ((u + 0x1000) & 0xFFFFE000) & 0xFFFFFFFF
    u = torch.where(u >= 2 ** 31, u - 2 ** 32, u).to(torch.int32)
    return u.view(torch.float32)


@pytest.mark.parametrize("N,K", [(16, 8), (16, 32), (64, 8), (64, 64), (128, 24), (256, 16), (32, 40)])
def test_umma_selftest_gemm(N, K):
    g = torch.Generator().manual_seed(N * 100 + K)
    A = tf32_rna(torch.randn(128, K, generator=g))
    B = tf32_rna(torch.randn(N, K, generator=g))
    out = torch.full((128, N), float("nan"), device=DEV)
    lib = _lib.load()
    Ad, Bd = A.to(DEV), B.to(DEV)          # keep the device tensors alive across the call
    _lib.check(lib.wctb_selftest_umma(out.data_ptr(), Ad.data_ptr(), Bd.data_ptr(), N, K,
                                      torch.cuda.current_stream().cuda_stream), "selftest_umma")
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * K ** 0.5 * 4, "max err %g" % err


CASES = [
    # H, W, cin, cout, epi
    (2, 2, 16, 16, 0), (16, 62, 16, 16, 0), (17, 63, 16, 16, 0), (40, 130, 16, 32, 0), (33, 70, 32, 32, 1),
    (34, 66, 64, 64, 0), (9, 9, 128, 64, 2), (40, 40, 16, 16, 2), (3, 3, 64, 128, 1), (65, 33, 32, 64, 0),
    (20, 200, 128, 128, 0), (30, 124, 128, 128, 2), (18, 62, 64, 64, 1), (7, 61, 32, 16, 2), (50, 125, 16, 16, 1),
    (12, 64, 256, 256, 0), (10, 70, 256, 512, 1), (9, 20, 512, 256, 2), (130, 250, 32, 32, 0),
]


@pytest.mark.parametrize("H,W,cin,cout,epi", CASES)
def test_conv_tf32_engine_exact_products(H, W, cin, cout, epi):
    assert ops.tf32_supported(cin, cout)
    g = torch.Generator().manual_seed(H * 1000 + W + cin + cout + epi)
    x = tf32_rna(torch.randn(1, cin, H, W, generator=g))
    w = tf32_rna(torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="reflect"), w.double(), b.double()))
    if epi == 1:
        ref = F.max_pool2d(ref, 2, 2)
    elif epi == 2:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    wp = ops.pack_weights(w.to(DEV), ops.ENGINE_TF32)
    y = ops.conv3x3_p4(ops.nchw_to_p4(x.to(DEV)), wp, b.to(DEV), cout, epi, False, ops.ENGINE_TF32)
    torch.cuda.synchronize()
    got = ops.p4_to_nchw(y).cpu().double()
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), "max err %g" % err
    # and the fp32 engine on the same (exactly representable) operands agrees
    y32 = ops.conv3x3_p4(ops.nchw_to_p4(x.to(DEV)), ops.pack_weights(w.to(DEV), ops.ENGINE_FP32), b.to(DEV), cout, epi,
                         False, ops.ENGINE_FP32)
    assert (ops.p4_to_nchw(y32).cpu().double() - got).abs().max().item() <= 4e-5 * max(1.0, ref.abs().max().item())


def test_conv_tf32_rounds_output_when_asked():
    g = torch.Generator().manual_seed(9)
    x = tf32_rna(torch.randn(1, 16, 20, 70, generator=g))
    w = tf32_rna(torch.randn(16, 16, 3, 3, generator=g) * 0.1)
    b = torch.randn(16, generator=g) * 0.1
    wp = ops.pack_weights(w.to(DEV), ops.ENGINE_TF32)
    y0 = ops.conv3x3_p4(ops.nchw_to_p4(x.to(DEV)), wp, b.to(DEV), 16, 0, False, ops.ENGINE_TF32)
    y1 = ops.conv3x3_p4(ops.nchw_to_p4(x.to(DEV)), wp, b.to(DEV), 16, 0, True, ops.ENGINE_TF32)
    assert torch.equal(tf32_rna(y0.cpu()), y1.cpu())
    assert torch.equal(y1.cpu(), tf32_rna(y1.cpu()))


def test_pack_weights_tf32_layout():
    cin, cout = 16, 32
    w = torch.randn(cout, cin, 3, 3, generator=torch.Generator().manual_seed(1))
    p = ops.pack_weights(w.to(DEV), ops.ENGINE_TF32).cpu().view(cin // 8, 9, 2, cout, 4)
    ref = tf32_rna(w).view(cout, cin // 8, 2, 4, 9).permute(1, 4, 2, 0, 3)      # [kg][tap][c][n][e]
    assert torch.equal(p, ref.contiguous())


@pytest.mark.parametrize("H,W,c1,n,epi", [(16, 62, 16, 16, 1), (2, 2, 16, 16, 0), (37, 131, 16, 16, 1), (50, 64, 16, 16, 0),
                                           (33, 63, 64, 64, 1), (18, 125, 64, 64, 0), (161, 200, 16, 16, 1)])
def test_fused_head_equals_two_layer_path(H, W, c1, n, epi):
    """conv11(FFMA)+conv12(tcgen05) fused == conv3x3_first -> conv3x3_p4(TF32): same arithmetic, same order."""
    g = torch.Generator().manual_seed(H + W + c1)
    x = torch.rand(1, 3, H, W, generator=g).to(DEV)
    w11 = (torch.randn(c1, 3, 3, 3, generator=g) * 0.3).to(DEV)
    b11 = (torch.randn(c1, generator=g) * 0.1).to(DEV)
    w12 = (torch.randn(n, c1, 3, 3, generator=g) * (2.0 / (9 * c1)) ** 0.5).to(DEV)
    b12 = (torch.randn(n, generator=g) * 0.1).to(DEV)
    p11 = ops.pack_weights(w11, ops.ENGINE_FP32)
    p12 = ops.pack_weights(w12, ops.ENGINE_TF32)
    mid = ops.conv3x3_first(x, p11, b11, c1, True)
    ref = ops.conv3x3_p4(mid, p12, b12, n, epi, False, ops.ENGINE_TF32)
    got = ops.conv_head(x, p11, b11, p12, b12, c1, n, epi, False)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("H,W,up", [(14, 60, False), (2, 2, False), (37, 131, False), (16, 64, False), (61, 200, False),
                                    (28, 120, True), (4, 4, True), (74, 262, True), (150, 64, True)])
def test_fused_tail_equals_two_layer_path(H, W, up):
    """[x2] conv12 + conv11 (both tcgen05, intermediate in smem) fused == conv3x3_p4(TF32, rounded) -> conv3x3_last
    with TF32-rounded conv11 weights (products exact, only fp32 accumulation order differs)."""
    g = torch.Generator().manual_seed(H * 7 + W)
    h, w = (H // 2, W // 2) if up else (H, W)
    x = tf32_rna(torch.randn(1, 16, h, w, generator=g).relu()).to(DEV)
    w12 = (torch.randn(16, 16, 3, 3, generator=g) * 0.12).to(DEV)
    b12 = (torch.randn(16, generator=g) * 0.1).to(DEV)
    w11 = tf32_rna(torch.randn(3, 16, 3, 3, generator=g) * 0.1).to(DEV)
    b11 = (torch.randn(3, generator=g) * 0.1 + 0.2).to(DEV)
    p12 = ops.pack_weights(w12, ops.ENGINE_TF32)
    p11 = ops.pack_weights(w11, ops.ENGINE_FP32)
    w11p = torch.zeros(16, 16, 3, 3, device=DEV)
    w11p[:3] = w11
    p11t = ops.pack_weights(w11p, ops.ENGINE_TF32)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    mid = ops.conv3x3_p4(ops.nchw_to_p4(xin), p12, b12, 16, 0, True, ops.ENGINE_TF32)     # TF32-rounded intermediate
    ref = ops.conv3x3_last(mid, p11, b11)
    got = ops.conv_tail(ops.nchw_to_p4(x), p12, b12, p11t, b11, up)
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (1, 3, H, W)
    assert (got - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("H,W,epi", [(16, 62, 1), (2, 2, 0), (37, 131, 1), (50, 64, 0), (161, 200, 1), (17, 63, 1), (33, 125, 0)])
def test_fused_head_tc_equals_two_layer_path(H, W, epi):
    """all-tensor-core head (conv11 via the LBO=16 tap-pair MMAs) == conv3x3_first -> conv3x3_p4(TF32) when image and
    conv11 weights are TF32-representable (exact products)."""
    g = torch.Generator().manual_seed(H * 3 + W)
    x = tf32_rna(torch.rand(1, 3, H, W, generator=g)).to(DEV)
    w11 = tf32_rna(torch.randn(16, 3, 3, 3, generator=g) * 0.3).to(DEV)
    b11 = (torch.randn(16, generator=g) * 0.1).to(DEV)
    w12 = (torch.randn(16, 16, 3, 3, generator=g) * 0.12).to(DEV)
    b12 = (torch.randn(16, generator=g) * 0.1).to(DEV)
    p11 = ops.pack_weights(w11, ops.ENGINE_FP32)
    p12 = ops.pack_weights(w12, ops.ENGINE_TF32)
    mid = ops.conv3x3_first(x, p11, b11, 16, True)
    ref = ops.conv3x3_p4(mid, p12, b12, 16, epi, False, ops.ENGINE_TF32)
    got = ops.conv_head_tc(x, ops.pack_head_tc_weights(w11), b11, p12, b12, epi, False)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    # conv11 differs only by fp32 accumulation order; a value that lands on a TF32 rounding boundary may flip one ulp
    # (2^-11 relative) in the intermediate, which conv12 damps: allow 2e-3 of the output scale on <0.1% of outputs
    d = (got - ref).abs()
    scale = max(1.0, ref.abs().max().item())
    assert d.max().item() <= 5e-3 * scale
    assert (d > 2e-5 * scale).float().mean().item() < 2e-2
