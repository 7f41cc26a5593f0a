#!/usr/bin/env python
"""Driver for ncu captures / CUDA-event timing of the fused h2 head and tail on a 3840x2160 image."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from collaborative_distillation_b200 import ops
H, W = 2160, 3840
g = torch.Generator().manual_seed(0)
x = torch.rand(1, 3, H, W, generator=g).cuda()
w11 = (torch.randn(16, 3, 3, 3, generator=g) * 40).cuda(); b11 = torch.randn(16, generator=g).cuda()
w12 = (torch.randn(16, 16, 3, 3, generator=g) * 0.1).cuda(); b12 = torch.randn(16, generator=g).cuda() * 0.1
w3 = (torch.randn(3, 16, 3, 3, generator=g) * 0.1).cuda(); b3 = torch.randn(3, generator=g).cuda() * 0.1
w11p, i11 = ops.pack_head_h2_w11(w11); w12p, i12 = ops.pack_dx_h2(w12); w3p, i3 = ops.pack_dx_h2(w3)
xh = ops.nchw_to_h8(torch.rand(16, H // 2, W // 2, generator=g).cuda())
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n = 1 if "--once" in sys.argv else 10
print("head %dx%d: %.3f ms" % (W, H, t(lambda: ops.conv_head_h2(x, w11p, i11, b11, w12p, i12, b12), n)))
print("tail %dx%d (ups): %.3f ms" % (W, H, t(lambda: ops.conv_tail_h2(xh, w12p, i12, b12, w3p, i3, b3, True), n)))
