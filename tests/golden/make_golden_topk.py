#!/usr/bin/env python
"""Golden vectors for the reference's eigenvalue-truncation knobs (build container only; reads /root/reference).

`util_wct.py:26-27` defines NumEigenValue = 30 and RatEigenValue = 0.25; their uses (`# k_c = NumEigenValue`,
`# k_c = int(cFSize[0] * RatEigenValue)` at :87-88 and the style twins at :113-114) are commented out in the shipped
file.  This script loads the reference source IN MEMORY, enables exactly those lines (one knob at a time, nothing else is
touched, nothing is written back), and runs the reference's own `whiten_and_color_torch` on the inputs already stored in
golden_wct.npz.  Output: golden_wct_topk.npz next to this file.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference, ref_args  # noqa: E402


def patched_wct_class(enable):
    """reference WCT class with the commented-out `k_c = ...` / `k_s = ...` lines for knob `enable` switched on"""
    import_reference()            # shims + chdir (module-level imports of util_wct.py resolve)
    src = open(os.path.join(REF, "PytorchWCT", "util_wct.py")).read()
    lines = {"num": ("# k_c = NumEigenValue", "# k_s = NumEigenValue"),
             "rat": ("# k_c = int(cFSize[0] * RatEigenValue)", "# k_s = int(sFSize[0] * RatEigenValue)")}[enable]
    for ln in lines:
        assert src.count(ln) >= 1, ln
        src = src.replace(ln, ln[2:])
    ns = {"__name__": "util_wct_patched"}
    exec(compile(src, "util_wct_patched.py", "exec"), ns)
    return ns


def main():
    g = np.load(os.path.join(HERE, "golden_wct.npz"))
    out = {}
    for knob in ("num", "rat"):
        ns = patched_wct_class(knob)
        wct = ns["WCT"](ref_args("16x"))
        for case, num in (("full_rank", 10), ("wide", 30), ("dead_channels", 12)):
            cF, sF = torch.from_numpy(g[case + ".cF"]).double(), torch.from_numpy(g[case + ".sF"]).double()
            if knob == "num":
                ns["NumEigenValue"] = num
                tag = "%s.num%d" % (case, num)
            else:
                tag = "%s.rat025" % case
            out[tag] = wct.whiten_and_color_torch(cF, sF).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_wct_topk.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
