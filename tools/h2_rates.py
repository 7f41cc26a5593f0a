#!/usr/bin/env python
"""tcgen05 microbenchmarks for the h2 engine: cycles per kind::f16 MMA (M=128, K=16) vs N / #accumulators, and
tcgen05.ld (32x32b.x16) throughput with 4 and 8 warps."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborative_distillation_b200 import _lib
lib = _lib.load()
out = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
print("kind::f16 M=128 K=16:  N nacc ctas  cycles/MMA")
for ctas in (1, 148):
    for N in (16, 32, 48, 64, 96, 128, 256):
        for na in sorted({1, min(8, 512 // N)}):
            iters = 200
            _lib.check(lib.wctb_debug_mma_rate_f16(out.data_ptr(), N, na, iters, ctas, st), "mma_rate_f16")
            torch.cuda.synchronize()
            c = out[:ctas].float().mean().item() / (iters * 4 * na)
            print("%3d  %d  %3d   %7.1f" % (N, na, ctas, c))
print("tcgen05.ld 32x32b.x16: nwarps ctas  cycles per x16 load per warp   bytes/clk/SM")
for ctas in (1, 148):
    for nw in (4, 8):
        iters, per = 200, 16
        out.zero_()
        _lib.check(lib.wctb_debug_ldtm_rate(out.data_ptr(), nw, per, iters, ctas, st), "ldtm_rate")
        torch.cuda.synchronize()
        cyc = out.view(-1, 8)[:ctas, :nw].float().max().item()
        per_load = cyc / (iters * per)
        print("%d  %3d   %7.1f   %7.1f" % (nw, ctas, per_load, nw * 2048.0 / per_load))
print("kind::f16 M=128 K=16, A start address off the 128-byte grid (dx taps: +0, +off, +2*off units of 16 B):  N nacc off  cycles/MMA")
for N in (32, 64, 128):
    na = min(8, 512 // N)
    for off in (0, 1, 2, 4, 8):
        iters = 200
        _lib.check(lib.wctb_debug_mma_rate_f16_off(out.data_ptr(), N, na, iters, 1, off, st), "mma_rate_f16_off")
        torch.cuda.synchronize()
        c = out[:1].float().mean().item() / (iters * 4 * na)
        print("%3d  %d  %d   %7.1f" % (N, na, off, c))
