import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "pending_hw: written without GPU access (round-1 GPU budget was spent); runs only with "
                            "WCTB_PENDING_HW=1 until its first green run on a B200 is recorded in profiles/")


def pytest_collection_modifyitems(config, items):
    import torch
    if os.environ.get("WCTB_PENDING_HW") != "1":
        pend = pytest.mark.skip(reason="pending first hardware validation (set WCTB_PENDING_HW=1)")
        for item in items:
            if "pending_hw" in item.keywords:
                item.add_marker(pend)
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
