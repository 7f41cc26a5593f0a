#!/usr/bin/env python
"""Recipe for oracle/_ref/ -- the reference's OWN implementation of the hot path, staged for the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (the product never imports it).  The reference is pure Python, so "building" it means
staging the files its WCT path imports, unmodified, from where they lie under /root/reference into oracle/_ref/ (listed
in .gitignore so reference sources never enter the history; NOT gpurun-ignored, so the directory travels to the GPU box
like the built .so files): PytorchWCT/util_wct.py, model/model_{cd,original,kd2sd}.py, utils.py and the shipped 16x
weights.  `oracle/ref_runner.py` imports them with the three shims of SURVEY 8(c) (no source edits) and is what
`bench.py --impl reference` times when the directory exists (cpu_baseline.kind = "reference"); otherwise the arm falls
back to the oracle port (kind = "port").
Run by __graft_entry__.build() whenever /root/reference is present (i.e. in the build container)."""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["PytorchWCT/util_wct.py", "utils.py", "model/__init__.py", "model/model_cd.py", "model/model_original.py",
         "model/model_kd2sd.py", "LICENSE"]
TREES = ["trained_models/wct_se_16x_new", "trained_models/wct_se_16x_new_sd"]


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/build_ref.py: %s not present (GPU box?) -- keeping whatever oracle/_ref already holds" % REF)
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(REF, f), d)
    for t in TREES:
        shutil.copytree(os.path.join(REF, t), os.path.join(DST, t))
    # the reference reaches `model/` and `utils.py` from PytorchWCT/ through symlinks (PytorchWCT/model -> ../model)
    os.symlink("../model", os.path.join(DST, "PytorchWCT", "model"))
    os.symlink("../utils.py", os.path.join(DST, "PytorchWCT", "utils.py"))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("oracle/_ref staged: %d files" % n)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
