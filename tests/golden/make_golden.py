#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference code (build container only).

Reads /root/reference (read-only), writes small .npz fixtures next to this file.
The GPU box has no /root/reference; tests only read the committed fixtures.

Shims (no reference source edits):
  1. stub `matplotlib` (utils.py:1 imports it; not installed here),
  2. `torch.utils.serialization.load_lua` stub (removed from torch >= 1.0),
  3. `WCT.transform` gets a pre-sized csF (util_wct.py:221 `csF.data.resize_` no longer
     resizes the caller's tensor on modern torch).

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    stub = tempfile.mkdtemp()
    os.makedirs(os.path.join(stub, "matplotlib"))
    with open(os.path.join(stub, "matplotlib", "__init__.py"), "w") as f:
        f.write("def use(*a, **k):\n    pass\n")
    open(os.path.join(stub, "matplotlib", "pyplot.py"), "w").close()
    sys.path.insert(0, stub)
    import torch.utils.serialization as S

    def _no_lua(*a, **k):
        raise RuntimeError("load_lua is not available")
    S.load_lua = _no_lua
    os.chdir(os.path.join(REF, "PytorchWCT"))
    sys.path.insert(0, ".")
    import util_wct  # noqa
    return util_wct


def ref_args(mode):
    a = SimpleNamespace(mode=mode, numpy=False)
    for k in range(1, 6):
        if mode == "16x":
            setattr(a, "e%d" % k, "../trained_models/wct_se_16x_new/%dSE.pth" % k)
            setattr(a, "d%d" % k, "../trained_models/wct_se_16x_new_sd/%dSD.pth" % k)
        else:
            setattr(a, "e%d" % k, None)
            setattr(a, "d%d" % k, None)
    return a


def conv_params(module):
    out = {}
    for k, v in module.state_dict().items():
        if "aux" in k:
            continue  # conv*_aux / aux* heads are never used by forward()
        out[k] = v.detach().cpu().numpy().astype(np.float32)
    return out


def run_reference(wct, content, style, alpha, stages=(5, 4, 3, 2, 1)):
    taps = {}
    img = content
    with torch.no_grad():
        for s in stages:
            enc, dec = getattr(wct, "e%d" % s), getattr(wct, "d%d" % s)
            sF = enc(style).squeeze(0)
            cF = enc(img).squeeze(0)
            csF = wct.transform(cF.clone(), sF.clone(), torch.empty(1, *cF.shape), alpha)
            img = dec(csF.clone())
            taps["cF%d" % s], taps["sF%d" % s], taps["csF%d" % s] = cF, sF, csF.squeeze(0)
            taps["img%d" % s] = img
    return img, taps


def main():
    util_wct = import_reference()
    torch.set_num_threads(8)

    # ---------------- 16x: shipped weights + 5-stage goldens ----------------
    wct = util_wct.WCT(ref_args("16x"))
    wct.eval()
    weights = {}
    for k in range(1, 6):
        for tag in ("e", "d"):
            for n, v in conv_params(getattr(wct, "%s%d" % (tag, k))).items():
                weights["%s%d.%s" % (tag, k, n)] = v
    np.savez(os.path.join(HERE, "weights_16x.npz"), **weights)

    g = torch.Generator().manual_seed(1234)
    content = torch.rand(1, 3, 84, 100, generator=g)        # not a multiple of 16: floor-pool bookkeeping
    style = torch.rand(1, 3, 72, 64, generator=g)
    # smooth the noise a little so features have natural-ish statistics
    content = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(content, (1, 1, 1, 1), mode="reflect"), 3, 1)
    style = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(style, (2, 2, 2, 2), mode="reflect"), 5, 1)
    out = {"content": content.numpy(), "style": style.numpy()}
    for alpha in (1.0, 0.6):
        img, taps = run_reference(wct, content, style, alpha)
        tag = "a%02d" % int(alpha * 10)
        for s in (5, 4, 3, 2, 1):
            out["%s.img%d" % (tag, s)] = taps["img%d" % s].numpy()
        for s in (5, 4, 3):
            for n in ("cF", "sF", "csF"):
                out["%s.%s%d" % (tag, n, s)] = taps["%s%d" % (n, s)].numpy()
        for s in (2, 1):  # large tensors: keep a strided sample + sums
            for n in ("cF", "sF", "csF"):
                t = taps["%s%d" % (n, s)]
                out["%s.%s%d.sub" % (tag, n, s)] = t[:, ::4, ::4].numpy()
                out["%s.%s%d.sum" % (tag, n, s)] = np.array([t.double().sum().item(), t.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, "golden_16x.npz"), **out)

    # ---------------- whiten_and_color unit goldens (torch + numpy variants) ----------------
    w = {}
    g = torch.Generator().manual_seed(7)
    cases = {"full_rank": (24, 900, 700), "wide": (64, 500, 300), "dead_channels": (32, 400, 350),
             "hw_lt_c": (48, 30, 40)}
    for name, (C, nc, ns) in cases.items():
        mixc = torch.randn(C, C, generator=g, dtype=torch.float64)
        mixs = torch.randn(C, C, generator=g, dtype=torch.float64)
        cF = torch.relu(mixc @ torch.randn(C, nc, generator=g, dtype=torch.float64) + 0.5).float()
        sF = torch.relu(mixs @ torch.randn(C, ns, generator=g, dtype=torch.float64) * 2 + 1.0).float()
        if name == "dead_channels":
            cF[[3, 17, 30]] = 0
            sF[[3, 17, 30]] = 0
        wct.args.numpy = False
        t = wct.whiten_and_color(cF.double(), sF.double())
        wct.args.numpy = True
        tn = wct.whiten_and_color(cF.double(), sF.double())
        wct.args.numpy = False
        w[name + ".cF"], w[name + ".sF"] = cF.numpy(), sF.numpy()
        w[name + ".out_torch"], w[name + ".out_numpy"] = t.numpy(), tn.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_wct.npz"), **w)

    # ---------------- original mode (random init under seed 0): stages 2,1 ----------------
    torch.manual_seed(0)
    wo = util_wct.WCT(ref_args("original"))
    wo.eval()
    ow = {}
    for k in (1, 2):
        for tag in ("e", "d"):
            for n, v in conv_params(getattr(wo, "%s%d" % (tag, k))).items():
                ow["%s%d.%s" % (tag, k, n)] = v
    g = torch.Generator().manual_seed(99)
    c64 = torch.rand(1, 3, 64, 64, generator=g)
    s64 = torch.rand(1, 3, 48, 80, generator=g)
    img, taps = run_reference(wo, c64, s64, 1.0, stages=(2, 1))
    oo = {"content": c64.numpy(), "style": s64.numpy(), "img2": taps["img2"].numpy(), "img1": taps["img1"].numpy(),
          "cF1": taps["cF1"].numpy(), "csF1": taps["csF1"].numpy()}
    # BASELINE.json configs[0]: 256x256 content+style, original mode, single stage-1 WCT
    torch.manual_seed(0)
    c256 = torch.rand(1, 3, 256, 256)
    s256 = torch.rand(1, 3, 256, 256)
    img, taps = run_reference(wo, c256, s256, 1.0, stages=(1,))
    oo["cfg1.img1.crop"] = img[:, :, 100:132, 60:92].numpy()
    oo["cfg1.img1.sum"] = np.array([img.double().sum().item(), img.double().abs().sum().item()])
    oo["cfg1.csF1.sub"] = taps["csF1"][:, ::16, ::16].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_original.npz"), **oo, **{"w." + k: v for k, v in ow.items()})
    for f in ("weights_16x.npz", "golden_16x.npz", "golden_wct.npz", "golden_original.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
