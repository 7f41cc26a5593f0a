"""Run the UNMODIFIED reference modules staged under oracle/_ref (see oracle/build_ref.py) on torch-cpu.

TEST / BENCH INFRASTRUCTURE ONLY.  Three shims, no source edits:
  1. a stub `matplotlib` package (utils.py:1 imports it; not installed here),
  2. `torch.utils.serialization.load_lua` stub (gone from torch >= 1.0; only reached for .t7 weights),
  3. `WCT.transform` is handed a pre-sized csF (util_wct.py:221 `csF.data.resize_` no longer resizes the caller's tensor).
The stage loop below is PytorchWCT/WCT.py:98-106,120-125 with the `.cuda()` calls dropped (WCT.py itself cannot be
imported without a GPU: module-level `.cuda()` at :97,:110)."""
import os
import sys
import tempfile
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_util_wct = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "PytorchWCT", "util_wct.py")) and os.path.isdir(os.path.join(REF, "trained_models"))


def _import():
    global _util_wct
    if _util_wct is not None:
        return _util_wct
    stub = tempfile.mkdtemp(prefix="wctb_mpl_stub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    with open(os.path.join(stub, "matplotlib", "__init__.py"), "w") as f:
        f.write("def use(*a, **k):\n    pass\n")
    open(os.path.join(stub, "matplotlib", "pyplot.py"), "w").close()
    try:
        import matplotlib  # noqa: F401
    except Exception:
        sys.path.insert(0, stub)
    import torch.utils.serialization as S

    def _no_lua(*a, **k):
        raise RuntimeError("load_lua is not available")
    S.load_lua = _no_lua
    cwd = os.getcwd()
    os.chdir(os.path.join(REF, "PytorchWCT"))          # relative weight paths + model/ utils.py symlinks
    sys.path.insert(0, os.path.join(REF, "PytorchWCT"))
    try:
        import importlib
        _util_wct = importlib.import_module("util_wct")
    finally:
        os.chdir(cwd)
    return _util_wct


def make_wct(mode="16x"):
    u = _import()
    a = SimpleNamespace(mode=mode, numpy=False)
    for k in range(1, 6):
        if mode == "16x":
            setattr(a, "e%d" % k, os.path.join(REF, "trained_models", "wct_se_16x_new", "%dSE.pth" % k))
            setattr(a, "d%d" % k, os.path.join(REF, "trained_models", "wct_se_16x_new_sd", "%dSD.pth" % k))
        else:
            setattr(a, "e%d" % k, None)
            setattr(a, "d%d" % k, None)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):    # "load model ... successfully" x 10
        w = u.WCT(a)
    w.eval()
    return w


@torch.no_grad()
def stylize(wct, content, style, alpha=1.0, stages=(5, 4, 3, 2, 1)):
    img = content
    for s in stages:                                   # WCT.py:121-125
        enc, dec = getattr(wct, "e%d" % s), getattr(wct, "d%d" % s)
        sF = enc(style).squeeze(0)                     # WCT.py:99-103
        cF = enc(img).squeeze(0)
        csF = wct.transform(cF, sF, torch.empty(1, *cF.shape), alpha)   # WCT.py:104 (pre-sized csF: shim 3)
        img = dec(csF)                                 # WCT.py:105
    return img
