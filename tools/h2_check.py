#!/usr/bin/env python
"""First-contact check + micro-timing of the h2 conv engine on a B200 (diagnostic tool, not a test).
Prints, per case, the max error against fp64 and WHERE it sits (row/col/channel blocks), then CUDA-event timings of the
cfg3 layer shapes for the h2 and TF32 engines."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
from collaborative_distillation_b200 import ops

DEV = "cuda"


def ref_conv(x, w, b, epi):
    r = F.relu(F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="reflect"), w.double(), b.double()))
    if epi == 1:
        r = F.max_pool2d(r, 2, 2)
    elif epi == 2:
        r = F.interpolate(r, scale_factor=2, mode="nearest")
    return r


def check(H, W, cin, cout, epi):
    g = torch.Generator().manual_seed(H + W + cin + cout)
    x = torch.randn(1, cin, H, W, generator=g) * 3
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = ref_conv(x, w, b, epi)
    wp, ws = ops.pack_weights_h2(w.to(DEV))
    y8, y4 = ops.conv3x3_h2(ops.nchw_to_h8(x.to(DEV)), wp, ws, b.to(DEV), cin, cout, epi, True, True)
    torch.cuda.synchronize()
    got = ops.p4_to_nchw(y4).cpu().double()
    d = (got - ref).abs()[0]
    scale = max(1.0, ref.abs().max().item())
    msg = "case H%d W%d %d->%d epi%d: max err %.3g (scale %.3g)" % (H, W, cin, cout, epi, d.max().item(), scale)
    if d.max().item() > 1e-4 * scale or not torch.isfinite(got).all():
        bad = (d > 1e-4 * scale) | ~torch.isfinite(got[0])
        cs = bad.flatten(1).any(1).nonzero().flatten().tolist()
        ys = bad.any(0).any(1).nonzero().flatten().tolist()
        xs = bad.any(0).any(0).nonzero().flatten().tolist()
        msg += "  BAD: %d els; channels %s rows %s cols %s" % (int(bad.sum()), cs[:20], ys[:24], xs[:24])
        msg += " got[0,:4,0,0]=%s ref=%s" % (got[0, :4, 0, 0].tolist(), ref[0, :4, 0, 0].tolist())
    print(msg, flush=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    for c in [(2, 2, 16, 16, 0), (16, 62, 16, 16, 0), (40, 130, 16, 16, 0), (40, 130, 16, 32, 1), (34, 66, 64, 64, 2),
              (20, 200, 128, 128, 0), (70, 260, 24, 16, 0), (300, 700, 16, 16, 1), (12, 64, 256, 256, 0)]:
        try:
            check(*c)
        except Exception as e:
            print("case %s raised %s" % (c, e), flush=True)
    if "--time" in sys.argv:
        shapes = [(2160, 3840, 16, 16, 1), (1080, 1920, 16, 32, 0), (1080, 1920, 32, 32, 1), (540, 960, 32, 64, 0), (540, 960, 64, 64, 0),
                  (270, 480, 64, 128, 0), (270, 480, 128, 128, 0), (1080, 1920, 32, 16, 2), (2160, 3840, 16, 16, 0)]
        for (H, W, cin, cout, epi) in shapes:
            x = torch.randn(1, cin, H, W, device=DEV)
            w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
            b = torch.zeros(cout, device=DEV)
            wp, ws = ops.pack_weights_h2(w)
            x8 = ops.nchw_to_h8(x)
            t_h2 = timeit(lambda: ops.conv3x3_h2(x8, wp, ws, b, cin, cout, epi, True, False))
            x4 = ops.nchw_to_p4(x, True)
            wt = ops.pack_weights(w, ops.ENGINE_TF32)
            t_tf = timeit(lambda: ops.conv3x3_p4(x4, wt, b, cout, epi, True, ops.ENGINE_TF32))
            fl = 2 * 9 * cin * cout * H * W
            ob = {0: 1, 1: 0.25, 2: 4}[epi]
            by = 4 * H * W * (cin + cout * ob)
            print("time %dx%d %d->%d epi%d: h2 %.3f ms (%.0f TFLOP/s, %.2f TB/s) | tf32 %.3f ms (%.0f TFLOP/s, %.2f TB/s)"
                  % (H, W, cin, cout, epi, t_h2, fl / t_h2 / 1e9, by / t_h2 / 1e9, t_tf, fl / t_tf / 1e9, by / t_tf / 1e9), flush=True)


if __name__ == "__main__":
    main()
