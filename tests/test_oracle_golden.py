"""The CPU oracle (oracle/wct_oracle.py) pinned against fixtures produced by the
reference's own code (tests/golden/make_golden.py).  No GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import wct_oracle as O


@pytest.fixture(scope="module")
def g16(golden_dir):
    return np.load(os.path.join(golden_dir, "golden_16x.npz"))


@pytest.fixture(scope="module")
def w16(golden_dir):
    return O.load_weights_npz(os.path.join(golden_dir, "weights_16x.npz"))


def test_plans_match_reference_shapes(w16):
    # every conv in the plan exists in the shipped state_dicts with the planned shape
    for s in range(1, 6):
        for item in O.encoder_plan("16x", s):
            if item != "P":
                n, cin, cout = item
                assert tuple(w16["e%d" % s][n + ".weight"].shape) == (cout, cin, 3, 3)
        for item in O.decoder_plan("16x", s):
            if item != "U":
                n, cin, cout = item
                assert tuple(w16["d%d" % s][n + ".weight"].shape) == (cout, cin, 3, 3)
    assert [O.feature_channels("16x", s) for s in range(1, 6)] == [24, 32, 64, 128, 128]
    assert [O.feature_channels("original", s) for s in range(1, 6)] == [64, 128, 256, 512, 512]


@pytest.mark.parametrize("alpha", [1.0, 0.6])
def test_five_stage_16x_matches_reference(g16, w16, alpha):
    content, style = torch.from_numpy(g16["content"]), torch.from_numpy(g16["style"])
    taps = {}
    torch.set_num_threads(8)
    img = O.stylize(w16, "16x", content, style, alpha=alpha, taps=taps)
    tag = "a%02d" % int(alpha * 10)
    # shape bookkeeping is bit exact: 84x100 -> 80x96 after stage 5 (floor pools), then stays
    assert tuple(taps["img5"].shape) == (1, 3, 80, 96)
    for s in (5, 4, 3, 2, 1):
        ref = g16["%s.img%d" % (tag, s)]
        got = taps["img%d" % s].numpy()
        assert got.shape == ref.shape
        # same torch build, same ops: expect ~bit-exact; allow LAPACK-order noise through the SVD
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    for s in (5, 4, 3):
        for n in ("cF", "sF", "csF"):
            ref = g16["%s.%s%d" % (tag, n, s)]
            np.testing.assert_allclose(taps["%s%d" % (n, s)].numpy(), ref, rtol=0, atol=2e-5 * max(1.0, np.abs(ref).max()))
    for s in (2, 1):
        for n in ("cF", "sF", "csF"):
            t = taps["%s%d" % (n, s)]
            np.testing.assert_allclose(t[:, ::4, ::4].numpy(), g16["%s.%s%d.sub" % (tag, n, s)], rtol=0,
                                       atol=2e-5 * max(1.0, float(t.abs().max())))
            sums = g16["%s.%s%d.sum" % (tag, n, s)]
            assert abs(t.double().abs().sum().item() - sums[1]) <= 1e-6 * sums[1]
    assert img.min() >= 0  # decoders end with ReLU; no clamp at the top (values > 1 allowed)


@pytest.mark.parametrize("case", ["full_rank", "wide", "dead_channels", "hw_lt_c"])
def test_whiten_and_color_matches_reference(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "golden_wct.npz"))
    cF, sF = torch.from_numpy(g[case + ".cF"]), torch.from_numpy(g[case + ".sF"])
    for variant, key in ((False, ".out_torch"), (True, ".out_numpy")):
        ref = g[case + key]
        got = O.whiten_and_color(cF, sF, numpy_variant=variant).numpy()
        scale = np.abs(ref).max()
        if case == "hw_lt_c" and not variant:
            # rank-deficient covariance without the +I regulariser: the reference multiplies SVD
            # noise eigenvalues (1e-16*lmax) by lambda^-1/2; only the range-space part is defined.
            assert np.isfinite(got).all()
            continue
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-7 * scale)


def test_original_mode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "golden_original.npz"))
    w = {}
    for k in g.files:
        if k.startswith("w."):
            net, name = k[2:].split(".", 1)
            w.setdefault(net, {})[name] = torch.from_numpy(g[k])
    content, style = torch.from_numpy(g["content"]), torch.from_numpy(g["style"])
    taps = {}
    O.stylize(w, "original", content, style, stages=(2, 1), taps=taps)
    np.testing.assert_allclose(taps["img2"].numpy(), g["img2"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(taps["img1"].numpy(), g["img1"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(taps["csF1"].numpy(), g["csF1"], rtol=0, atol=2e-5 * np.abs(g["csF1"]).max())
    # BASELINE.json configs[0]: 256x256 / 256x256, original mode, stage 1 only, torch-cpu
    torch.manual_seed(0)
    c256 = torch.rand(1, 3, 256, 256)
    s256 = torch.rand(1, 3, 256, 256)
    taps = {}
    img = O.stylize(w, "original", c256, s256, stages=(1,), taps=taps)
    np.testing.assert_allclose(img[:, :, 100:132, 60:92].numpy(), g["cfg1.img1.crop"], rtol=0, atol=2e-5)
    assert abs(img.double().abs().sum().item() - g["cfg1.img1.sum"][1]) <= 1e-6 * g["cfg1.img1.sum"][1]


TOPK_CASES = [("full_rank", "num10", 10), ("wide", "num30", 30), ("dead_channels", "num12", 12),
              ("full_rank", "rat025", int(24 * 0.25)), ("wide", "rat025", int(64 * 0.25)), ("dead_channels", "rat025", int(32 * 0.25))]


@pytest.mark.parametrize("case,tag,keep", TOPK_CASES)
def test_eigenvalue_truncation_matches_reference(golden_dir, case, tag, keep):
    """util_wct.py:26-27,87-88,113-114 (NumEigenValue / RatEigenValue): fixtures from the reference with its own
    commented-out lines enabled in memory (tests/golden/make_golden_topk.py)."""
    g = np.load(os.path.join(golden_dir, "golden_wct.npz"))
    ref = np.load(os.path.join(golden_dir, "golden_wct_topk.npz"))["%s.%s" % (case, tag)]
    cF, sF = torch.from_numpy(g[case + ".cF"]), torch.from_numpy(g[case + ".sF"])
    got = O.whiten_and_color(cF, sF, keep=keep).numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-7 * np.abs(ref).max())
    # and the knob really truncates: the coloured, centred output has rank <= keep
    centred = got - got.mean(1, keepdims=True)
    sv = np.linalg.svd(centred, compute_uv=False)
    assert (sv > 1e-8 * sv[0]).sum() <= keep
