"""CPU oracle for the WCT stylization hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-cpu / fp64 restatement of the reference algorithm
(MingSun-Tse/Collaborative-Distillation, PytorchWCT/WCT.py + util_wct.py +
model/model_{cd,original,kd2sd}.py).  It is NOT part of the product: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it, and only as the checker (or as the timed CPU
baseline).  The product path (`collaborative_distillation_b200`) never imports
it and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or known-answer
fixtures for this path, so the oracle is pinned against the
outputs of the reference's own code imported in the build container:
`tests/golden/make_golden.py` drives the unmodified reference classes and
writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file
against them (runs everywhere, no GPU, no /root/reference needed).

Arithmetic that lives in third-party dependencies of the reference: torch
(requirements.txt:3 pins torch==0.4.1; conv2d / ReflectionPad2d / MaxPool2d /
UpsamplingNearest2d / mm semantics are unchanged in torch 2.11 used here) and
LAPACK via torch.svd (util_wct.py:74,100).  Only (V, E) of the SVD of the
symmetric PSD covariance are consumed, so any symmetric eigensolver is
equivalent up to sign/rotation inside degenerate eigenspaces.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------
# Architecture tables.
# VGG-19 prefix: blocks of 3x3 convs separated by 2x2 max-pools.
#   reference: model/model_cd.py:688-702 (16x encoder 5), 246-258 (16x decoder 5),
#              model/model_original.py:434-446 / 539-551 (unpruned),
#              model/model_cd.py:324 (encoder 1 of the 16x family is 3->24).
# ---------------------------------------------------------------------------
VGG_LAYERS = ["conv11", "conv12", "P", "conv21", "conv22", "P",
              "conv31", "conv32", "conv33", "conv34", "P",
              "conv41", "conv42", "conv43", "conv44", "P", "conv51"]

WIDTHS = {
    "original": {1: 64, 2: 128, 3: 256, 4: 512, 5: 512},
    "16x": {1: 16, 2: 32, 3: 64, 4: 128, 5: 128},
}
WIDTHS["16x_kd2sd"] = WIDTHS["16x"]


def encoder_plan(mode: str, stage: int):
    """[(name, cin, cout) | 'P'] for encoder `stage` (ends at conv{stage}1).

    model_cd.py:346-349,403-409,485-494,589-603,724-743; model_original.py:36-39 ...
    """
    w = dict(WIDTHS[mode])
    if mode != "original" and stage == 1:
        w[1] = 24  # model_cd.py:324
    plan, cin = [], 3
    for item in VGG_LAYERS:
        if item == "P":
            plan.append("P")
            continue
        cout = w[int(item[4])]
        plan.append((item, cin, cout))
        cin = cout
        if item == "conv%d1" % stage:
            break
    return plan


def decoder_plan(mode: str, stage: int):
    """[(name, cin, cout) | 'U'] for decoder `stage`: mirror of the encoder.

    model_cd.py:83-85,117-122,159-167,211-224,276-294; model_original.py:581-599.
    conv{k}1 maps width[k] -> width[k-1]; the final conv11 maps width[1] -> 3.
    """
    enc = encoder_plan(mode, stage)
    plan = []
    for item in reversed(enc):
        if item == "P":
            plan.append("U")
            continue
        name, cin, cout = item
        plan.append((name, cout, cin))
    return plan


def feature_channels(mode: str, stage: int) -> int:
    return [p for p in encoder_plan(mode, stage) if p != "P"][-1][2]


# ---------------------------------------------------------------------------
# Encoder / decoder forward (fp32, torch-cpu ops = the reference's own backend)
# ---------------------------------------------------------------------------
def _conv3x3_reflect_relu(x, w, b):
    # nn.ReflectionPad2d((1,1,1,1)) -> nn.Conv2d(k=3,s=1,p=0) -> ReLU, e.g. model_cd.py:725-726
    return F.relu(F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b))


def encoder_forward(params: dict, mode: str, stage: int, x: torch.Tensor) -> torch.Tensor:
    """x: [1,3,H,W] fp32 in [0,1] -> [1,C,h,w].  params: {"conv0.weight": ..., ...}."""
    y = F.conv2d(x, params["conv0.weight"], params["conv0.bias"])  # 1x1, model_cd.py:725
    for item in encoder_plan(mode, stage):
        if item == "P":
            y = F.max_pool2d(y, 2, 2)  # floor mode, model_cd.py:709
        else:
            n = item[0]
            y = _conv3x3_reflect_relu(y, params[n + ".weight"], params[n + ".bias"])
    return y


def decoder_forward(params: dict, mode: str, stage: int, y: torch.Tensor) -> torch.Tensor:
    """y: [1,C,h,w] -> [1,3,H,W]; ReLU after the last conv too (model_cd.py:293)."""
    for item in decoder_plan(mode, stage):
        if item == "U":
            y = F.interpolate(y, scale_factor=2, mode="nearest")  # UpsamplingNearest2d, model_cd.py:261
        else:
            n = item[0]
            y = _conv3x3_reflect_relu(y, params[n + ".weight"], params[n + ".bias"])
    return y


# ---------------------------------------------------------------------------
# whiten_and_color / transform  (fp64; util_wct.py:62-131, 134-202, 210-223)
# ---------------------------------------------------------------------------
EIGEN_VALUE_THRE = 1e-100  # util_wct.py:25


def _rank(e):
    # util_wct.py:82-86 / 108-112: first index whose eigenvalue is below the threshold
    k = e.numel()
    for i in range(e.numel()):
        if e[i] < EIGEN_VALUE_THRE:
            k = i
            break
    return k


def whiten_and_color(cF: torch.Tensor, sF: torch.Tensor, numpy_variant: bool = False, keep: int | None = None) -> torch.Tensor:
    """cF [C,HWc], sF [C,HWs] (any float dtype; computed in fp64) -> [C,HWc] fp64.

    numpy_variant=True reproduces whiten_and_color_np's `+ eye(C)` on the content
    covariance (util_wct.py:143); the torch path has it commented out (util_wct.py:70).
    """
    cF = cF.double()
    sF = sF.double()
    C, n_c = cF.shape
    c_mean = cF.mean(1, keepdim=True)                      # util_wct.py:68
    cFc = cF - c_mean                                      # :69
    c_cov = (cFc @ cFc.t()) / (n_c - 1)                    # :70
    if numpy_variant:
        c_cov = c_cov + torch.eye(C, dtype=torch.float64)  # :143
    _, c_e, c_vh = torch.linalg.svd(c_cov)                 # :74 (only E, V are consumed)
    c_v = c_vh.t()
    k_c = _rank(c_e)

    n_s = sF.shape[1]
    s_mean = sF.mean(1, keepdim=True)                      # :94
    sFc = sF - s_mean                                      # :95
    s_cov = (sFc @ sFc.t()) / (n_s - 1)                    # :96
    _, s_e, s_vh = torch.linalg.svd(s_cov)                 # :100
    s_v = s_vh.t()
    k_s = _rank(s_e)
    if keep is not None:                                   # :87-88, :113-114 (commented out in the reference):
        k_c, k_s = min(k_c, int(keep)), min(k_s, int(keep))   # k = NumEigenValue or int(C * RatEigenValue)

    c_d = c_e[:k_c].pow(-0.5)                              # :117
    whiten = (c_v[:, :k_c] * c_d) @ c_v[:, :k_c].t() @ cFc  # :118-120
    s_d = s_e[:k_s].pow(0.5)                               # :124
    target = (s_v[:, :k_s] * s_d) @ s_v[:, :k_s].t() @ whiten  # :125
    return target + s_mean                                 # :126


def transform(cF: torch.Tensor, sF: torch.Tensor, alpha: float, numpy_variant: bool = False) -> torch.Tensor:
    """cF [C,H,W], sF [C,H1,W1] fp32 -> csF [1,C,H,W] fp32 (util_wct.py:210-223)."""
    C = cF.shape[0]
    cD = cF.double()
    t = whiten_and_color(cD.reshape(C, -1), sF.double().reshape(C, -1), numpy_variant).view_as(cD)
    cs = alpha * t + (1.0 - alpha) * cD                    # :219 (blend with the uncentred cF)
    return cs.float().unsqueeze(0)                         # :220


# ---------------------------------------------------------------------------
# Driver: 5-stage coarse-to-fine loop (PytorchWCT/WCT.py:98-106, 120-125)
# ---------------------------------------------------------------------------
def style_transfer_stage(weights: dict, mode: str, stage: int, content: torch.Tensor, style: torch.Tensor,
                         alpha: float = 1.0, numpy_variant: bool = False, taps: dict | None = None) -> torch.Tensor:
    """One `styleTransfer(wct.eK, wct.dK, cImg, sImg, csF)` call (WCT.py:98-106)."""
    with torch.no_grad():
        sF = encoder_forward(weights["e%d" % stage], mode, stage, style).squeeze(0)
        cF = encoder_forward(weights["e%d" % stage], mode, stage, content).squeeze(0)
        csF = transform(cF, sF, alpha, numpy_variant)
        out = decoder_forward(weights["d%d" % stage], mode, stage, csF)
    if taps is not None:
        taps["cF%d" % stage], taps["sF%d" % stage], taps["csF%d" % stage] = cF, sF, csF.squeeze(0)
    return out


def stylize(weights: dict, mode: str, content: torch.Tensor, style: torch.Tensor, alpha: float = 1.0,
            num_run: int = 1, stages=(5, 4, 3, 2, 1), numpy_variant: bool = False, taps: dict | None = None):
    """content/style: [1,3,H,W] fp32.  Returns the stage-1 output image (not clamped)."""
    img = content
    for _ in range(num_run):                               # WCT.py:120
        for s in stages:                                   # WCT.py:121-125
            img = style_transfer_stage(weights, mode, s, img, style, alpha, numpy_variant, taps)
            if taps is not None:
                taps["img%d" % s] = img
    return img


# ---------------------------------------------------------------------------
# Weights helpers (random init with the reference's fixed conv0, or from npz)
# ---------------------------------------------------------------------------
def random_weights(mode: str, seed: int = 0, stages=(1, 2, 3, 4, 5), scale: float = 1.0) -> dict:
    """He-style random weights of the right shapes (NOT the reference's nn.Conv2d default
    init; used for oracle-vs-CUDA parity where any weights do).  conv0 is the fixed
    RGB->BGR*255-mean affine of model_original.py:427-433."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for s in stages:
        e = {"conv0.weight": torch.tensor([[0, 0, 255.], [0, 255., 0], [255., 0, 0]]).view(3, 3, 1, 1),
             "conv0.bias": torch.tensor([-103.939, -116.779, -123.68])}
        for item in encoder_plan(mode, s):
            if item == "P":
                continue
            n, cin, cout = item
            std = scale * (2.0 / (9 * cin)) ** 0.5
            if n == "conv11":
                std = std / 64.0  # inputs are ~[-124, 151] after conv0
            e[n + ".weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
            e[n + ".bias"] = torch.randn(cout, generator=g) * 0.05
        d = {}
        for item in decoder_plan(mode, s):
            if item == "U":
                continue
            n, cin, cout = item
            std = scale * (2.0 / (9 * cin)) ** 0.5
            d[n + ".weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
            d[n + ".bias"] = torch.randn(cout, generator=g) * 0.05 + (0.3 if cout == 3 else 0.0)
        out["e%d" % s], out["d%d" % s] = e, d
    return out


def load_weights_npz(path: str) -> dict:
    """npz with keys 'e5.conv11.weight' ... -> {'e5': {'conv11.weight': tensor}, ...}."""
    z = np.load(path)
    out: dict = {}
    for k in z.files:
        net, name = k.split(".", 1)
        out.setdefault(net, {})[name] = torch.from_numpy(z[k])
    return out
