#!/usr/bin/env python
"""Time the content-side whitening solvers on the covariances of a cfg3 step (B200): Jacobi eigensolver at several
early-stop thresholds vs the Newton-Schulz path (wctb_whiten_ns), and the error of each W against LAPACK."""
import os, sys
from types import SimpleNamespace
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import ops
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
P.set_precision("h2")
w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
P.weights.load_npz_into(w, os.path.join(root, "tests", "golden", "weights_16x.npz"))
w = w.cuda()
g = torch.Generator().manual_seed(0)
c = torch.rand(1, 3, 2160, 3840, generator=g).cuda()


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for s in (5, 4, 3, 2, 1):
    x4 = getattr(w, "e%d" % s).forward_p4(c)
    C = x4.shape[0] * 4
    n = float(x4.shape[1] * x4.shape[2])
    mean = ops.channel_sum(x4) / n
    gram = ops.centered_gram(x4, mean).view(1, C, C)
    S = (gram[0] / (n - 1)).cpu()
    lam, V = torch.linalg.eigh(S)
    keep = lam > 1e-7 * lam.max()
    ref = (V[:, keep] * lam[keep].rsqrt()) @ V[:, keep].t()
    line = "stage %d C=%d live=%d cond=%.1e |" % (s, C, int(keep.sum()), (lam.max() / lam[keep].min()).item())
    for early in (3e-6, 1e-4, 1e-3, 1e-2):
        t = timeit(lambda: ops.eigh_jacobi(gram, [1.0 / (n - 1)], early_stop=early))
        ev, evec, sw = ops.eigh_jacobi(gram, [1.0 / (n - 1)], early_stop=early, return_sweeps=True)
        ev, evec = ev[0].cpu(), evec[0].cpu()
        k = ev > 1e-7 * ev.max()
        Wj = (evec[k].t() * ev[k].rsqrt()) @ evec[k]
        line += " jac(%.0e) %.3f ms sw=%d err=%.1e |" % (early, t, int(sw[0]), ((Wj - ref).abs().max() / ref.abs().max()).item())
    if C <= 128:
        t = timeit(lambda: ops.whiten_ns(gram[0], 1.0 / (n - 1)))
        Wn, info = ops.whiten_ns(gram[0], 1.0 / (n - 1), return_info=True)
        line += " ns %.3f ms it=%s err=%.1e" % (t, info.cpu().tolist(), ((Wn.cpu() - ref).abs().max() / ref.abs().max()).item())
    print(line, flush=True)
    del x4
