#!/usr/bin/env python
"""Run each hot conv kernel once on a cfg3-sized tensor (for `ncu -k regex:...` captures)."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collaborative_distillation_b200 as P  # noqa: E402

P.set_precision("tf32")
w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
P.weights.load_npz_into(w, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "weights_16x.npz"))
w = w.cuda()
x = torch.rand(1, 3, 2160, 3840, device="cuda")
for _ in range(2):
    f = w.e2.forward_p4(x)        # conv_head<16,16,pool> + conv_umma<32,0>
    img = w.d2.forward_p4(f)      # conv_umma<16,0>(32->16) + conv_tail<UPSRC>
torch.cuda.synchronize()
print("ok", tuple(f.shape), tuple(img.shape))
