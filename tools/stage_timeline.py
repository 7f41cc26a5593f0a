#!/usr/bin/env python
"""CUDA-event timeline of the content (critical) path of one cfg3 stylization, per stage and phase."""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collaborative_distillation_b200 as P
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P.set_precision(sys.argv[1] if len(sys.argv) > 1 else "h2")
print("precision", P.get_precision())
w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
P.weights.load_npz_into(w, os.path.join(root, "tests", "golden", "weights_16x.npz"))
w = w.cuda()
g = torch.Generator().manual_seed(0)
c = torch.rand(1, 3, 2160, 3840, generator=g).cuda(); s = torch.rand(1, 3, 2000, 2000, generator=g).cuda()
for _ in range(3):
    w.stylize(c, s)
torch.cuda.synchronize()
w.timeline = []
w.stylize(c, s)
torch.cuda.synchronize()
tl = w.timeline
tot = {}
print("stage  enc    stats   eig    matrix+dec")
for st in (5, 4, 3, 2, 1):
    ev = {n: e for (s_, n, e) in tl if s_ == st}
    d = [ev["start"].elapsed_time(ev["enc"]), ev["enc"].elapsed_time(ev["stats"]), ev["stats"].elapsed_time(ev["eig"]), ev["eig"].elapsed_time(ev["dec"])]
    print("  %d   %6.3f %6.3f %6.3f %6.3f" % (st, *d))
    for k, v in zip(("enc", "stats", "eig", "dec"), d):
        tot[k] = tot.get(k, 0) + v
print("total ", " ".join("%s=%.3f" % kv for kv in tot.items()), " sum=%.3f ms" % sum(tot.values()))
