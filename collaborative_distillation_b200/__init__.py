"""collaborative_distillation_b200 -- B200-native WCT stylization hot path
(drop-in for MingSun-Tse/Collaborative-Distillation's PytorchWCT/WCT.py + util_wct.py + model/*)."""
from . import arch, image_io, nets, ops, weights  # noqa: F401
from ._lib import WctbError, load  # noqa: F401
from .nets import get_precision, set_precision  # noqa: F401
from .util_wct import WCT  # noqa: F401
from .pipeline import StylizePipeline  # noqa: F401

__all__ = ["WCT", "StylizePipeline", "nets", "ops", "arch", "image_io", "set_precision", "get_precision", "WctbError", "load"]
