#!/usr/bin/env python
"""Phase timing of the fused head kernel (clock64 stamps of 8 CTAs in tile row 10)."""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import _lib
P.set_precision("tf32")
w = P.WCT(SimpleNamespace(mode="16x", numpy=False)).cuda()
x = torch.rand(1, 3, 2160, 3840, device="cuda")
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
w.e2.forward_p4(x); torch.cuda.synchronize()
_lib.check(_lib.load().wctb_debug_set_trace(buf.data_ptr()), "trace")
w.e2.forward_p4(x); torch.cuda.synchronize()
_lib.load().wctb_debug_set_trace(None)
t = buf.cpu().view(8, 16)[:, :8]
names = ["entry", "setup done", "P0 img tile", "M1 conv11 mma", "E1 convert", "M2 conv12 mma", "E2 pool epi", "exit sync"]
d = (t[:, 1:] - t[:, :-1]).float()
print("phase durations (cycles), 8 CTAs:")
for i in range(7):
    print("  %-16s mean %8.0f  min %8.0f  max %8.0f" % (names[i + 1], d[:, i].mean(), d[:, i].min(), d[:, i].max()))
print("  total            mean %8.0f" % (t[:, 7] - t[:, 0]).float().mean())
