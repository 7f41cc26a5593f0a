"""Reference import path `from model.model_kd2sd import ...` -> the B200-native classes (same names and signatures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from collaborative_distillation_b200 import nets as _n  # noqa: E402

for _k, _v in vars(_n).items():
    if isinstance(_v, type) and issubclass(_v, _n._Net) and not _k.startswith("_"):
        globals()[_k] = _v
