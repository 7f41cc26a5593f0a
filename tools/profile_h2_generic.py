#!/usr/bin/env python
"""Driver for ncu captures / CUDA-event timing of the generic h2 conv kernel at the cfg3 layer shapes.
usage: profile_h2_generic.py [--once] [cin:cout:H:W:epi ...]   (default: the four heaviest cfg3 shapes)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from collaborative_distillation_b200 import ops
specs = [a for a in sys.argv[1:] if ":" in a] or ["64:64:540:960:0", "128:128:270:480:0", "32:32:1080:1920:1", "16:32:1080:1920:0", "32:16:1080:1920:0"]
once = "--once" in sys.argv
noflush = "--noflush" in sys.argv      # back-to-back launches on the same tensors: whatever fits the 126 MB L2 stays resident
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sp in specs:
    cin, cout, H, W, epi = (int(v) for v in sp.split(":"))
    x = ops.nchw_to_h8(torch.rand(cin, H, W, generator=g).cuda())
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (1.0 / (9 * cin) ** 0.5)).cuda(); b = torch.randn(cout, generator=g).cuda() * 0.1
    wp, ws = ops.pack_weights_h2(w)
    fn = lambda: ops.conv3x3_h2(x, wp, ws, b, cin, cout, epi)
    n = 1 if once else 10
    for _ in range(0 if once else 2):
        fn()
    ts = []
    for _ in range(n):
        if not noflush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * 9 * cin * cout * H * W
    Ho, Wo = (H // 2, W // 2) if epi == 1 else ((2 * H, 2 * W) if epi == 2 else (H, W))
    by = 4.0 * (cin * H * W + cout * Ho * Wo)
    print(("[no flush] " if noflush else "") + "conv_h2 %d->%d %dx%d epi%d: %.3f ms  %.1f TFLOP/s  %.0f GB/s" % (cin, cout, W, H, epi, ms, fl / ms * 1e-9, by / ms * 1e-6))
