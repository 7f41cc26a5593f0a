#!/usr/bin/env python
"""CPU probe (test infrastructure; imports oracle/): error of operand-rounded convolutions chained through the 5-stage
16x pipeline, against the fp32 oracle, at BASELINE cfg2 (1024^2 / 512^2) for the two input families the parity tests use
(torch.rand seed 0 = what bench.py feeds; a natural pair).  `bits` = significand bits kept (incl. the implicit one):
11 = TF32 / fp16, 8 = bf16, 16 = bf16 hi+lo, 22 = fp16 hi+lo.
usage: python tools/precision_probe.py [--size 1024 512] [--natural path_c path_s]"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import wct_oracle as O


def round_bits(x, bits):
    if bits >= 24:
        return x
    drop = 24 - bits
    u = x.contiguous().view(torch.int32)
    half = 1 << (drop - 1)
    u = (u + half) & ~((1 << drop) - 1)          # round-to-nearest, ties away (cvt.rna)
    return u.view(torch.float32)


def run(weights, c, s, bits, wbits=None):
    wbits = bits if wbits is None else wbits
    orig = O._conv3x3_reflect_relu
    def patched(x, w, b):
        return orig(round_bits(x, bits), round_bits(w, wbits), b)
    O._conv3x3_reflect_relu = patched
    try:
        taps = {}
        out = O.stylize(weights, "16x", c, s, taps=taps)
    finally:
        O._conv3x3_reflect_relu = orig
    return out, taps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[1024, 512])
    ap.add_argument("--natural", nargs=2, default=None)
    ap.add_argument("--bits", type=int, nargs="*", default=[11, 8, 16])
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    W = O.load_weights_npz(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_16x.npz"))
    cs, ss = a.size
    pairs = {}
    torch.manual_seed(0)
    pairs["rand"] = (torch.rand(1, 3, cs, cs), torch.rand(1, 3, ss, ss))
    if a.natural:
        from PIL import Image
        def load(p, n):
            im = Image.open(p).convert("RGB").resize((n, n), Image.BILINEAR)
            return torch.from_numpy(np.asarray(im)).permute(2, 0, 1)[None].float() / 255
        pairs["natural"] = (load(a.natural[0], cs), load(a.natural[1], ss))
    for name, (c, s) in pairs.items():
        t0 = time.time()
        ref, rt = run(W, c, s, 24)
        print(f"[{name}] oracle fp32 {time.time()-t0:.1f}s  out range [{ref.min():.3f},{ref.max():.3f}]", flush=True)
        for bits in a.bits:
            out, tp = run(W, c, s, bits)
            line = f"[{name}] bits={bits:2d}  final rms {((out-ref)**2).mean().sqrt():.3e} max {(out-ref).abs().max():.3e} | per-stage img rms:"
            for st in (5, 4, 3, 2, 1):
                line += f" {((tp['img%d'%st]-rt['img%d'%st])**2).mean().sqrt():.2e}"
            print(line, flush=True)


if __name__ == "__main__":
    main()
