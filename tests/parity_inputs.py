"""Seeded input pairs shared by the parity tests, bench.py's accuracy leg and tools/: `rand` is exactly what bench.py feeds
(torch.rand, seed 0); `natural` is the reference's sample pair (tests/golden/natural_pair.npz) resized like
`--content_size/--style_size` (transforms.Resize on a PIL image = bilinear, antialiased)."""
import io
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rand_pair(hc, wc, hs, ws, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, 3, hc, wc, generator=g), torch.rand(1, 3, hs, ws, generator=g)


def natural_pair(hc, wc, hs, ws):
    from PIL import Image
    z = np.load(os.path.join(GOLDEN, "natural_pair.npz"))

    def load(key, h, w):
        im = Image.open(io.BytesIO(z[key].tobytes())).convert("RGB")
        if im.size != (w, h):
            im = im.resize((w, h), Image.BILINEAR)
        return torch.from_numpy(np.asarray(im).copy()).permute(2, 0, 1)[None].float() / 255

    return load("content_jpg", hc, wc), load("style_jpg", hs, ws)


def pair(kind, hc, wc, hs, ws):
    return rand_pair(hc, wc, hs, ws) if kind == "rand" else natural_pair(hc, wc, hs, ws)
