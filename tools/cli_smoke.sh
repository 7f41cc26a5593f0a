#!/bin/bash
# end-to-end CLI check on the GPU box: synthesise two content and two style images, run the reference-compatible CLI
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
mkdir -p $T/content $T/style $T/out $T/tm/wct_se_16x_new $T/tm/wct_se_16x_new_sd
python - <<PY
import numpy as np, torch, os
from PIL import Image
rng = np.random.default_rng(0)
for d, n, hw in (("content", "a.jpg", (212, 300)), ("content", "b.png", (180, 256)), ("style", "s1.jpg", (160, 200)), ("style", "s2.png", (200, 144))):
    Image.fromarray((rng.random((*hw, 3)) * 255).astype("uint8")).resize((hw[1], hw[0])).save(os.path.join("$T", d, n))
import sys; sys.path.insert(0, ".")
from types import SimpleNamespace
import collaborative_distillation_b200 as P
w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
P.weights.load_npz_into(w, "tests/golden/weights_16x.npz")
for k in range(1, 6):
    torch.save({"epoch": 20, "model": getattr(w, "e%d" % k).state_dict()}, os.path.join("$T", "tm", "wct_se_16x_new", "%dSE.pth" % k))
    torch.save({"epoch": 20, "model": getattr(w, "d%d" % k).state_dict()}, os.path.join("$T", "tm", "wct_se_16x_new_sd", "%dSD.pth" % k))
PY
cd PytorchWCT
python WCT.py --debug --mode 16x --contentPath $T/content --stylePath $T/style --outf $T/out --weights_root $T/tm --log_mark t 2>&1 | tail -4
ls $T/out
