#!/usr/bin/env python
"""Eigensolver diagnostics on the benchmark workload: live rank, sweeps and time per solve of both solver variants."""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import ops
P.set_precision("tf32")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
P.weights.load_npz_into(w, os.path.join(root, "tests", "golden", "weights_16x.npz"))
w = w.cuda()
g = torch.Generator().manual_seed(0)
x = torch.rand(1, 3, 2160, 3840, generator=g).cuda()
for s in (5, 4, 3, 2, 1):
    f = getattr(w, "e%d" % s).forward_p4(x)
    C = f.shape[0] * 4
    n = float(f.shape[1] * f.shape[2])
    gram = torch.zeros(1, C, C, device="cuda", dtype=torch.float64)
    mean = w._moments(f, (0, f.shape[1], 0, f.shape[2]), n, gram[0])
    live = int((gram[0].diagonal() > 0).sum())
    line = "stage %d C=%3d live=%3d" % (s, C, live)
    for variant, name in ((1, "legacy"), (0, "cholesky")):
        ops.set_eigh_variant(variant)
        for _ in range(2):
            ev, evec, sw = ops.eigh_jacobi(gram, [1.0 / (n - 1)], return_sweeps=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.eigh_jacobi(gram, [1.0 / (n - 1)])
        e1.record(); torch.cuda.synchronize()
        line += "  | %s sweeps=%2d %.3f ms" % (name, int(sw[0]), e0.elapsed_time(e1) / 5)
    evs = ev[0].sort(descending=True).values
    print(line + "  | lmax=%.3g  l[live-1]/lmax=%.2e" % (evs[0].item(), (evs[live - 1] / evs[0]).item()))
    if 64 < C <= 128:
        ops.set_eigh_variant(0)
        pr = ops.eigh_profile(gram, [1.0 / (n - 1)])
        tot = float(pr["load"] + pr["cholesky"] + pr["sweeps_phase"])
        print("   profile (clock64 of thread 0, k=%d, %d sweeps): load %.0f, cholesky %.0f, sweeps %.0f cycles (%.1f%% / %.1f%% / %.1f%%)" % (
            pr["k"], pr["sweeps"], pr["load"], pr["cholesky"], pr["sweeps_phase"], 100 * pr["load"] / tot,
            100 * pr["cholesky"] / tot, 100 * pr["sweeps_phase"] / tot))
lat, thr = ops.dp_rate()
print("fp64 pipe: %.1f cycles per dependent DFMA, %.1f DFMA/clk/SM with 16 warps x 8 independent chains" % (lat, thr))
