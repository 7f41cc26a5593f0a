#!/usr/bin/env python
"""Which TF32 pieces contribute how much error?  5-stage stylize vs the CPU oracle on (a) the smoke case (uniform-noise
96x128 / 64x80) and (b) a smoothed 256x320 / 200x240 case, toggling the optional tensor-core pieces.
Lives under tests/ because it uses the oracle as the checker (the oracle is test infrastructure only)."""
import os, sys
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import nets
from oracle import wct_oracle as O
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wpath = os.path.join(root, "tests", "golden", "weights_16x.npz")
ow = O.load_weights_npz(wpath)
g = torch.Generator().manual_seed(0)
cases = {"noise 96x128": (torch.rand(1, 3, 96, 128, generator=g), torch.rand(1, 3, 64, 80, generator=g))}
c2, s2 = torch.rand(1, 3, 256, 320, generator=g), torch.rand(1, 3, 200, 240, generator=g)
sm = lambda t, k: torch.nn.functional.avg_pool2d(torch.nn.functional.pad(t, (k // 2,) * 4, mode="reflect"), k, 1)
cases["smooth 256x320"] = (sm(c2, 5), sm(s2, 7))
torch.set_num_threads(16)
refs = {k: O.stylize(ow, "16x", c, s) for k, (c, s) in cases.items()}


def run(label, precision="tf32", head_tc=True, fuse_head=True, fuse_tail=True, fold=True, fast_gram=True):
    P.set_precision(precision)
    nets.HEAD_TC, nets.FUSE_HEAD, nets.FUSE_TAIL = head_tc, fuse_head, fuse_tail
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    P.weights.load_npz_into(w, wpath)
    w = w.cuda()
    w.fold_into_decoder = fold
    w.use_graph = False
    if not fast_gram:
        import collaborative_distillation_b200.ops as ops
        orig = ops.centered_gram
        ops.centered_gram = lambda *a, **k: orig(*a, **{**k, "fast": False})
    out = []
    for k, (c, s) in cases.items():
        y = w.stylize(c.cuda(), s.cuda()).cpu()
        d = y - refs[k]
        out.append("%s rms %.2e max %.2e" % (k, d.pow(2).mean().sqrt().item(), d.abs().max().item()))
    if not fast_gram:
        ops.centered_gram = orig
    print("%-44s %s" % (label, " | ".join(out)))


run("fp32 engine", precision="fp32")
run("tf32 all on")
run("tf32, FFMA conv11 in head (HEAD_TC off)", head_tc=False)
run("tf32, no fused head", fuse_head=False)
run("tf32, no fused tail (fp32 last conv)", fuse_tail=False)
run("tf32, no head/tail fusion", fuse_head=False, fuse_tail=False)
run("tf32, no fold (fp32 apply)", fold=False)
run("tf32, fp64 gram", fast_gram=False)
run("tf32, no fusion, no fold, fp64 gram", fuse_head=False, fuse_tail=False, fold=False, fast_gram=False)
