#!/usr/bin/env python
"""A/B of the two fp32-product Gram kernels (wctb_debug_set_gram_variant) on the cfg3 / cfg4 stage-1 and stage-2 feature
maps: CUDA-event time per launch (L2 flushed between launches) and agreement of the results.  Appends to
gpurun_out/gram_ab.txt.   python tools/gram_ab.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collaborative_distillation_b200 import ops  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "gram_ab.txt"), "a")


def log(s):
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (C, H, W) in [(24, 2160, 3840), (32, 1080, 1920), (24, 2000, 2000), (24, 4096, 10240)]:
    x = torch.rand(C // 4, H, W, 4, device="cuda") * 3
    mean = (ops.channel_sum(x) / (H * W))
    res = {}
    variants = ((1, "staged"), (2, "regs/L1"), (0, "regs/cp.async ring")) + (((3, "ring, peeled"), (4, "ring, 2 px/thread")) if "--peeled" in sys.argv else ())
    for variant, _ in variants:
        ops.set_gram_variant(variant)
        ts = []
        for i in range(6):
            g = torch.zeros(C, C, device="cuda", dtype=torch.float64)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.centered_gram(x, mean, out=g, fast=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[variant] = (sorted(ts[1:])[len(ts[1:]) // 2], g)
    ops.set_gram_variant(0)
    g64 = ops.centered_gram(x, mean)
    torch.cuda.synchronize()
    err = {v: ((res[v][1] - g64).abs().max() / g64.abs().max()).item() for v in res}
    gb = C * H * W * 4 / 1e9
    log("C=%d %dx%d (%.0f MB): " % (C, H, W, gb * 1e3) + " | ".join(
        "%s %.3f ms (%.0f GB/s, err %.1e)" % (name, res[v][0], gb / res[v][0] * 1e3, err[v])
        for v, name in variants))
    del x
