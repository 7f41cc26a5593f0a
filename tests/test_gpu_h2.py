"""h2 engine (fp16 hi/lo operand pairs on tcgen05.mma.kind::f16, csrc/conv_h2.cu): layout, weight packing, convolution
parity against an fp64 reference on ARBITRARY fp32 operands (nothing is pre-rounded: the point of the engine is that it
does not need that), and the path-level parity on BASELINE configs[1] / [2] against the CPU oracle.
Tolerances are written next to each assertion; measured values are in profiles/r02_*."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import ops
from oracle import wct_oracle as O
import parity_inputs  # tests/parity_inputs.py (pytest puts tests/ on sys.path)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def split_ref(x: torch.Tensor):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi, lo


# ------------------------------------------------------------------ layout + packing (bit-exact)
@pytest.mark.parametrize("shape", [(8, 5, 7), (24, 33, 65), (128, 16, 18), (3, 9, 11)])
def test_h8_layout_roundtrip(shape):
    C, H, W = shape
    x = torch.randn(C, H, W, device=DEV) * 37.0
    h8 = ops.nchw_to_h8(x)
    C8 = (C + 7) // 8
    assert tuple(h8.shape) == (C8, 2, H, W, 8) and h8.dtype == torch.float16
    xp = torch.zeros(C8 * 8, H, W, device=DEV)
    xp[:C] = x
    hi, lo = split_ref(xp)
    ref = torch.stack([hi.view(C8, 8, H, W).permute(0, 2, 3, 1), lo.view(C8, 8, H, W).permute(0, 2, 3, 1)], dim=1)
    assert torch.equal(h8, ref.contiguous())
    back = ops.h8_to_nchw(h8, C).squeeze(0)
    assert (back - x).abs().max().item() <= 2.0 ** -21 * x.abs().max().item()        # 22 significand bits
    if C % 4 == 0:
        assert torch.equal(ops.p4_to_h8(ops.nchw_to_p4(x)), h8)


@pytest.mark.parametrize("cin,cout", [(16, 16), (24, 16), (32, 64), (128, 256)])
def test_pack_weights_h2_layout_and_scale(cin, cout):
    g = torch.Generator().manual_seed(cin + cout)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    packed, ws = ops.pack_weights_h2(w.to(DEV))
    ws = ws.cpu()
    m = w.abs().max().item()
    s = ws[2].item()
    assert 512.0 <= m * s < 1024.0 and np.log2(s) == round(np.log2(s)) and ws[1].item() * s == 1.0
    N = cout if cout <= 128 else 128
    nkg = (cin + 15) // 16
    p = packed.cpu().view(cout // N, nkg, 9, 2, 2 * N, 8)
    wp = torch.zeros(cout, nkg * 16, 3, 3)
    wp[:, :cin] = w * s
    hi, lo = split_ref(wp)
    # [nb][kg][tap][c][r][e] <- w[nb*N + r][kg*16 + c*8 + e][tap]
    def lay(t):
        return t.view(cout // N, N, nkg, 2, 8, 9).permute(0, 2, 5, 3, 1, 4)
    assert torch.equal(p[..., :N, :], lay(hi).contiguous())
    assert torch.equal(p[..., N:, :], lay(lo).contiguous())


# ------------------------------------------------------------------ convolution parity vs fp64
CASES = [
    # H, W, cin, cout, epi
    (2, 2, 16, 16, 0), (16, 62, 16, 16, 0), (17, 63, 16, 16, 0), (40, 130, 16, 32, 0), (33, 70, 32, 32, 1),
    (34, 66, 64, 64, 0), (9, 9, 128, 64, 2), (40, 40, 16, 16, 2), (3, 3, 64, 128, 1), (65, 33, 32, 64, 0),
    (20, 200, 128, 128, 0), (30, 124, 128, 128, 2), (18, 62, 64, 64, 1), (7, 61, 32, 16, 2), (50, 125, 16, 16, 1),
    (12, 64, 256, 256, 0), (10, 70, 256, 512, 1), (9, 20, 512, 256, 2), (130, 250, 32, 32, 0), (70, 260, 24, 16, 0),
    (300, 700, 16, 16, 1),          # many interior (tensor-map TMA) tiles per CTA: exercises the persistent pipeline
    (200, 400, 64, 32, 2),
]


def _conv_ref(x, w, b, epi):
    ref = F.relu(F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="reflect"), w.double(), b.double()))
    if epi == 1:
        ref = F.max_pool2d(ref, 2, 2)
    elif epi == 2:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    return ref


@pytest.mark.parametrize("H,W,cin,cout,epi", CASES)
def test_conv_h2_vs_fp64(H, W, cin, cout, epi):
    assert ops.h2_supported(cin, cout)
    g = torch.Generator().manual_seed(H * 1000 + W + cin + cout + epi)
    x = torch.randn(1, cin, H, W, generator=g) * 3.0
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = _conv_ref(x, w, b, epi)
    wp, ws = ops.pack_weights_h2(w.to(DEV))
    y8, y4 = ops.conv3x3_h2(ops.nchw_to_h8(x.to(DEV)), wp, ws, b.to(DEV), cin, cout, epi, out_h8=True, out_p4=True)
    torch.cuda.synchronize()
    got4 = ops.p4_to_nchw(y4).cpu().double()
    got8 = ops.h8_to_nchw(y8, cout).cpu().double()
    assert got4.shape == ref.shape and got8.shape == ref.shape
    scale = max(1.0, ref.abs().max().item())
    err = (got4 - ref).abs().max().item()
    # operands carry 22 bits, accumulation is fp32 over K = 9*cin terms: same class as an fp32 engine
    # measured on B200: 6e-7 (cin 16) ... 9e-6 (cin 256) of the output scale: the tensor core accumulates with truncation, so
    # the error grows ~linearly in K = 9*cin (same behaviour as the TF32 engine and cuDNN tensor-op paths)
    assert err <= 2e-6 * scale * max(1.0, cin / 16.0), "fp32 output: max err %g (scale %g)" % (err, scale)
    assert (got8 - got4).abs().max().item() <= 2.0 ** -21 * scale      # the H8 copy is the same numbers, split


@pytest.mark.parametrize("H,W,cin", [(2, 2, 16), (17, 40, 24), (33, 131, 64), (80, 300, 16)])
def test_conv_h2_last_layer_nchw3(H, W, cin):
    g = torch.Generator().manual_seed(H + W + cin)
    x = torch.randn(1, cin, H, W, generator=g)
    w = torch.randn(3, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(3, generator=g) * 0.1
    ref = _conv_ref(x, w, b, 0)
    wpad = torch.zeros(16, cin, 3, 3)
    wpad[:3] = w
    bpad = torch.zeros(16)
    bpad[:3] = b
    wp, ws = ops.pack_weights_h2(wpad.to(DEV))
    _, img = ops.conv3x3_h2(ops.nchw_to_h8(x.to(DEV)), wp, ws, bpad.to(DEV), cin, 16, ops.EPI_NCHW3)
    assert tuple(img.shape) == (1, 3, H, W)
    assert (img.cpu().double() - ref).abs().max().item() <= 2e-6 * max(1.0, cin / 16.0) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("H,W,cout", [(2, 2, 16), (9, 13, 24), (37, 70, 16), (64, 64, 64), (17, 129, 24)])
def test_conv_first_h2(H, W, cout):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(1, 3, H, W, generator=g)
    w = torch.randn(cout, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(cout, generator=g) * 0.1
    wp = ops.pack_weights(w.to(DEV), ops.ENGINE_FP32)
    y8, y4 = ops.conv3x3_first_h2(x.to(DEV), wp, b.to(DEV), cout, out_h8=True, out_p4=True)
    ref = ops.conv3x3_first(x.to(DEV), wp, b.to(DEV), cout, False)
    assert torch.equal(y4, ref)                                        # same FFMA order as the fp32 first-layer kernel
    assert torch.equal(y8, ops.p4_to_h8(ref))


def test_conv_h2_large_values_saturate_not_nan():
    """fp16 hi saturates at 65504 instead of becoming inf (activations of the shipped nets stay below ~300)"""
    x = torch.full((1, 16, 8, 8), 1.0e5)
    h8 = ops.nchw_to_h8(x.to(DEV))
    back = ops.h8_to_nchw(h8, 16)
    assert torch.isfinite(back).all() and back.max().item() >= 65504.0


# ------------------------------------------------------------------ fused head (conv11 + conv12 + pool, dx-stacked, block-pipelined)
@pytest.mark.parametrize("H,W", [(2, 2), (4, 6), (16, 28), (33, 29), (37, 131), (64, 56), (100, 30), (161, 200), (300, 700),
                                 (31, 57), (32, 58), (65, 85)])
def test_fused_head_h2_vs_fp64_and_two_layer_path(H, W):
    g = torch.Generator().manual_seed(H * 7 + W)
    x = torch.rand(1, 3, H, W, generator=g)
    w11 = torch.randn(16, 3, 3, 3, generator=g) * 40.0             # conv0-folded magnitudes (x255)
    b11 = torch.randn(16, generator=g) * 5.0
    w12 = torch.randn(16, 16, 3, 3, generator=g) * (2.0 / 144) ** 0.5
    b12 = torch.randn(16, generator=g) * 0.1
    mid = F.relu(F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="reflect"), w11.double(), b11.double()))
    ref = F.max_pool2d(F.relu(F.conv2d(F.pad(mid, (1, 1, 1, 1), mode="reflect"), w12.double(), b12.double())), 2, 2)
    w11p, is11 = ops.pack_head_h2_w11(w11.to(DEV))
    w12p, is12 = ops.pack_dx_h2(w12.to(DEV))
    y = ops.conv_head_h2(x.to(DEV), w11p, is11, b11.to(DEV), w12p, is12, b12.to(DEV))
    torch.cuda.synchronize()
    got = ops.h8_to_nchw(y, 16).cpu().double()
    assert got.shape == ref.shape
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    assert err <= 1e-5 * scale, "fused head vs fp64: max err %g (scale %g)" % (err, scale)
    # the unfused h2 path (fp32 FFMA first layer -> generic h2 conv + pool) computes the same thing
    y1, _ = ops.conv3x3_first_h2(x.to(DEV), ops.pack_weights(w11.to(DEV), ops.ENGINE_FP32), b11.to(DEV), 16)
    wp, ws = ops.pack_weights_h2(w12.to(DEV))
    y2, _ = ops.conv3x3_h2(y1, wp, ws, b12.to(DEV), 16, 16, ops.EPI_POOL2)
    assert (ops.h8_to_nchw(y2, 16).cpu().double() - got).abs().max().item() <= 1e-5 * scale


@pytest.mark.parametrize("H,W,ups", [(2, 2, 0), (4, 6, 1), (16, 28, 0), (34, 30, 1), (37, 131, 0), (64, 56, 1), (100, 30, 1),
                                     (162, 200, 1), (300, 700, 1), (31, 57, 0), (32, 58, 1), (66, 86, 1), (17, 29, 0)])
def test_fused_tail_h2_vs_fp64_and_unfused_path(H, W, ups):
    g = torch.Generator().manual_seed(H * 5 + W + ups)
    h, w = (H // 2, W // 2) if ups else (H, W)
    x = torch.randn(1, 16, h, w, generator=g).abs() * 3.0
    w12 = torch.randn(16, 16, 3, 3, generator=g) * (2.0 / 144) ** 0.5
    b12 = torch.randn(16, generator=g) * 0.1
    w11 = torch.randn(3, 16, 3, 3, generator=g) * (2.0 / 144) ** 0.5
    b11 = torch.randn(3, generator=g) * 0.1 + 0.3
    xin = F.interpolate(x.double(), scale_factor=2, mode="nearest") if ups else x.double()
    mid = F.relu(F.conv2d(F.pad(xin, (1, 1, 1, 1), mode="reflect"), w12.double(), b12.double()))
    ref = F.relu(F.conv2d(F.pad(mid, (1, 1, 1, 1), mode="reflect"), w11.double(), b11.double()))
    w12p, is12 = ops.pack_dx_h2(w12.to(DEV))
    w11p, is11 = ops.pack_dx_h2(w11.to(DEV))
    x8 = ops.nchw_to_h8(x.to(DEV))
    got = ops.conv_tail_h2(x8, w12p, is12, b12.to(DEV), w11p, is11, b11.to(DEV), bool(ups))
    torch.cuda.synchronize()
    assert tuple(got.shape) == tuple(ref.shape) == (1, 3, H, W)
    scale = max(1.0, ref.abs().max().item())
    err = (got.cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * scale, "fused tail vs fp64: max err %g (scale %g)" % (err, scale)


# ------------------------------------------------------------------ modules with the shipped weights
def _wct16(precision):
    P.set_precision(precision)
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    return w.to(DEV)


def test_h2_encoders_decoders_vs_oracle_shipped_weights(golden_dir):
    wpath = os.path.join(golden_dir, "weights_16x.npz")
    w = _wct16("h2")
    P.weights.load_npz_into(w, wpath)
    ow = O.load_weights_npz(wpath)
    x = torch.rand(1, 3, 75, 110, generator=torch.Generator().manual_seed(5))
    for s in range(1, 6):
        ref = O.encoder_forward(ow["e%d" % s], "16x", s, x)
        got = getattr(w, "e%d" % s)(x.to(DEV)).cpu()
        assert got.shape == ref.shape
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel <= 5e-5, "encoder %d rel %g" % (s, rel)             # measured <= 2.2e-5 (fp32 CUDA-core engine: <= 5e-6)
        assert (got - ref).abs().max().item() <= 5e-5 * ref.abs().max().item()
        refd = O.decoder_forward(ow["d%d" % s], "16x", s, ref)
        gotd = getattr(w, "d%d" % s)(ref.to(DEV)).cpu()
        assert gotd.shape == refd.shape
        rel = ((gotd - refd).norm() / refd.norm()).item()
        assert rel <= 5e-5, "decoder %d rel %g" % (s, rel)


@pytest.mark.parametrize("alpha,fold", [(1.0, True), (0.6, True), (1.0, False)])
def test_h2_five_stage_vs_reference_golden(golden_dir, alpha, fold):
    g = np.load(os.path.join(golden_dir, "golden_16x.npz"))
    w = _wct16("h2")
    P.weights.load_npz_into(w, os.path.join(golden_dir, "weights_16x.npz"))
    w.fold_into_decoder = fold
    content, style = torch.from_numpy(g["content"]).to(DEV), torch.from_numpy(g["style"]).to(DEV)
    tag = "a%02d" % int(alpha * 10)
    img = content
    for s in (5, 4, 3, 2, 1):
        img = w.style_transfer_stage(s, img, style, alpha)
        ref = torch.from_numpy(g["%s.img%d" % (tag, s)])
        assert tuple(img.shape) == tuple(ref.shape)
        d = img.cpu() - ref
        rms, mx = d.pow(2).mean().sqrt().item(), d.abs().max().item()
        print("h2 stage %d alpha %.1f fold %s: rms %.3g max %.3g" % (s, alpha, fold, rms, mx))
        assert rms <= 5e-5 and mx <= 5e-4, "stage %d rms %g max %g" % (s, rms, mx)   # the fp32 engine's bounds


# ------------------------------------------------------------------ --mode 16x_kd2sd (model_kd2sd.py:52-70, WCT.py:60-70)
@pytest.mark.parametrize("precision,rms_tol,max_tol", [("h2", 2e-4, 4e-3), ("fp32", 2e-4, 4e-3)])
def test_mode_16x_kd2sd_five_stage_vs_oracle(golden_dir, precision, rms_tol, max_tol):
    """SmallDecoder{1..5}_16x_aux: the kd2sd decoders are the 16x conv stacks plus aux heads that forward() never uses
    (model_kd2sd.py:52-70).  Their weights are not shipped: shipped 16x encoders + seeded random decoders, whose state_dict
    (INCLUDING the aux* tensors) is loaded strictly like the reference's .pth; 5 stages against the CPU oracle."""
    P.set_precision(precision)
    w = P.WCT(SimpleNamespace(mode="16x_kd2sd", numpy=False))
    P.weights.load_npz_into(w, os.path.join(golden_dir, "weights_16x.npz"), stages=())    # nothing yet
    z = np.load(os.path.join(golden_dir, "weights_16x.npz"))
    ow = O.random_weights("16x_kd2sd", seed=21, scale=0.6)
    with torch.no_grad():
        for s in range(1, 6):
            enc, dec = getattr(w, "e%d" % s), getattr(w, "d%d" % s)
            for k in z.files:                                                   # shipped encoder weights
                net, name = k.split(".", 1)
                if net == "e%d" % s:
                    mod, attr = name.rsplit(".", 1)
                    getattr(getattr(enc, mod), attr).copy_(torch.from_numpy(z[k]))
                    ow[net][name] = torch.from_numpy(z[k])
            sd = dict(ow["d%d" % s])
            for name, cin, cout in P.arch.DECODER_AUX_KD2SD[s]:                 # aux heads exist in the .pth, unused by forward()
                sd[name + ".weight"] = torch.randn(cout, cin, 1, 1)
                sd[name + ".bias"] = torch.randn(cout)
            assert type(dec).__name__ == "SmallDecoder%d_16x_aux" % s
            dec.load_state_dict(sd, strict=True)
    w = w.to(DEV)
    content, style = parity_inputs.pair("natural", 192, 256, 160, 160)
    ref = O.stylize(ow, "16x_kd2sd", content, style)
    out = w.stylize(content.to(DEV), style.to(DEV)).cpu()
    P.set_precision("h2")
    assert out.shape == ref.shape
    d = out - ref
    rms, mx = d.pow(2).mean().sqrt().item(), d.abs().max().item()
    scale = max(1.0, ref.abs().max().item())
    print("16x_kd2sd %s: rms %.3g max %.3g (range [%.2f, %.2f])" % (precision, rms, mx, ref.min().item(), ref.max().item()))
    assert rms <= rms_tol * scale and mx <= max_tol * scale


# ------------------------------------------------------------------ BASELINE configs: the benched path vs the CPU oracle
# SURVEY 8(d) precision contract for the end-to-end image (5 stages, range [0, ~1.5]): 3e-3 RMS / 6e-2 max.
CONTRACT_RMS, CONTRACT_MAX = 3e-3, 6e-2


def _staged(w, content, style):
    img, out = content, {}
    for s in (5, 4, 3, 2, 1):
        img = w.style_transfer_stage(s, img, style, 1.0)
        out[s] = img
    return out


@pytest.mark.parametrize("kind", ["rand", "natural"])
def test_cfg2_path_parity_vs_oracle(golden_dir, kind):
    """BASELINE configs[1]: 1024x1024 content / 512x512 style, --mode 16x, all 5 stages, both engines that can be benched,
    on the image family bench.py feeds (torch.rand seed 0) and on a natural pair; error printed per stage."""
    wpath = os.path.join(golden_dir, "weights_16x.npz")
    content, style = parity_inputs.pair(kind, 1024, 1024, 512, 512)
    taps = {}
    O.stylize(O.load_weights_npz(wpath), "16x", content, style, taps=taps)
    res = {}
    for precision in ("h2", "tf32", "fp32"):
        w = _wct16(precision)
        P.weights.load_npz_into(w, wpath)
        imgs = _staged(w, content.to(DEV), style.to(DEV))
        # and the public entry point the bench times (two streams + CUDA graph) gives the same image
        pub = w.stylize(content.to(DEV), style.to(DEV))
        torch.cuda.synchronize()
        line = "cfg2 %-7s %-5s" % (kind, precision)
        for s in (5, 4, 3, 2, 1):
            d = imgs[s].cpu() - taps["img%d" % s]
            line += "  s%d %.2e/%.2e" % (s, d.pow(2).mean().sqrt().item(), d.abs().max().item())
        d = imgs[1].cpu() - taps["img1"]
        res[precision] = (d.pow(2).mean().sqrt().item(), d.abs().max().item())
        dp = pub.cpu() - taps["img1"]
        line += "  | stylize() %.2e/%.2e" % (dp.pow(2).mean().sqrt().item(), dp.abs().max().item())
        print(line)
        if precision == "h2":
            assert dp.pow(2).mean().sqrt().item() <= CONTRACT_RMS and dp.abs().max().item() <= CONTRACT_MAX
    P.set_precision("h2")
    # the default (benched) engine meets the contract with a wide margin; the fp32 CUDA-core engine is the yardstick
    assert res["h2"][0] <= CONTRACT_RMS / 10 and res["h2"][1] <= CONTRACT_MAX / 10, res
    assert res["h2"][0] <= 4 * res["fp32"][0] + 1e-6, res
    # single-pass TF32 is recorded, not required: it misses the contract on noise-like inputs
    assert res["tf32"][0] <= 0.2, res


def test_cfg3_path_parity_vs_oracle(golden_dir):
    """BASELINE configs[2] (3840x2160 / 2000x2000, what bench.py times): the default engine against the CPU oracle on the
    bench inputs themselves (~10 s of host time)."""
    wpath = os.path.join(golden_dir, "weights_16x.npz")
    content, style = parity_inputs.pair("rand", 2160, 3840, 2000, 2000)
    ref = O.stylize(O.load_weights_npz(wpath), "16x", content, style)
    w = _wct16("h2")
    P.weights.load_npz_into(w, wpath)
    out = w.stylize(content.to(DEV), style.to(DEV)).cpu()
    d = out - ref
    rms, mx = d.pow(2).mean().sqrt().item(), d.abs().max().item()
    print("cfg3 rand h2: rms %.3g max %.3g (range [%.2f, %.2f])" % (rms, mx, ref.min().item(), ref.max().item()))
    assert tuple(out.shape) == tuple(ref.shape) == (1, 3, 2160, 3840)
    assert rms <= CONTRACT_RMS / 10 and mx <= CONTRACT_MAX / 10
