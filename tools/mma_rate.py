#!/usr/bin/env python
"""tcgen05.mma issue-rate microbenchmark: cycles per MMA (M=128, K=8 tf32) vs N, operand layout, #accumulators."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborative_distillation_b200 import _lib
lib = _lib.load()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("N  layout nacc ctas  cycles/MMA")
for ctas in (1, 148):
    for N in (16, 64, 128, 256):
        for layout in (0, 1, 2):
            nacc = min(8, 512 // N)
            for na in sorted({1, nacc}):
                iters = 200
                _lib.check(lib.wctb_debug_mma_rate(out.data_ptr(), N, layout, na, iters, ctas, st), "mma_rate")
                torch.cuda.synchronize()
                c = out[:ctas].float().mean().item() / (iters * 4 * na)
                print("%3d  %d     %d    %3d   %7.1f" % (N, layout, na, ctas, c))
