"""Weight containers: npz <-> WCT modules, and deterministic synthetic weights for benchmarks."""
from __future__ import annotations

import numpy as np
import torch

from . import arch


def load_npz_into(wct, path: str, stages=(1, 2, 3, 4, 5)):
    """npz keys 'e5.conv11.weight' ... (tests/golden/weights_16x.npz) -> wct.e*/d* parameters (aux heads untouched)."""
    z = np.load(path)
    with torch.no_grad():
        for k in z.files:
            net, name = k.split(".", 1)
            if int(net[1]) not in stages:
                continue
            mod, attr = name.rsplit(".", 1)
            getattr(getattr(getattr(wct, net), mod), attr).copy_(torch.from_numpy(z[k]))
    return wct


def state_as_oracle_dict(wct, stages=(1, 2, 3, 4, 5)) -> dict:
    """{'e5': {'conv11.weight': cpu tensor, ...}, ...} in the layout oracle/wct_oracle.py consumes (tests only)."""
    out = {}
    for s in stages:
        for tag in ("e", "d"):
            net = getattr(wct, "%s%d" % (tag, s))
            out["%s%d" % (tag, s)] = {k: v.detach().cpu().clone() for k, v in net.state_dict().items() if "aux" not in k}
    return out


def synthetic_init_(wct, seed: int = 0):
    """Deterministic He-style weights with a VGG-normalised-like conv0 (benchmarks with no checkpoint:
    activations keep O(1..100) scale through all layers so the covariance spectra are well conditioned)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for s in range(1, 6):
            e, d = getattr(wct, "e%d" % s), getattr(wct, "d%d" % s)
            e.conv0.weight.copy_(torch.tensor([[0., 0, 255], [0, 255., 0], [255., 0, 0]]).view(3, 3, 1, 1))
            e.conv0.bias.copy_(torch.tensor([-103.939, -116.779, -123.68]))
            for net in (e, d):
                for L in net.layers:
                    conv = getattr(net, L["name"])
                    std = (2.0 / (9 * L["cin"])) ** 0.5
                    if net is e and L["name"] == "conv11":
                        std /= 64.0
                    conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * std)
                    conv.bias.copy_(torch.randn(conv.bias.shape, generator=g) * 0.05 + (0.3 if L["cout"] == 3 else 0.0))
    return wct
