"""`from util_wct import WCT` keeps working (reference: PytorchWCT/util_wct.py:30)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborative_distillation_b200.util_wct import WCT, EigenValueThre, TAU  # noqa: E402,F401
