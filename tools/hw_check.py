#!/usr/bin/env python
"""One-process hardware check for a short GPU slot: smoke() of the default path first (protects the round-end run),
then the tests still marked `pending_hw`.  Everything is appended to gpurun_out/hw_check.log as it happens.

    tests/native/validate_io gpurun_out/native_io.txt; python tools/hw_check.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "hw_check.log"), "a")


def log(msg):
    line = "[%7.1fs] %s" % (time.time() - T0, msg)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()


T0 = time.time()
log("start")
import torch  # noqa: E402

log("torch imported, cuda=%s" % torch.cuda.is_available())
import __graft_entry__ as G  # noqa: E402

try:
    G.smoke()
    log("smoke: OK")
except Exception as e:  # noqa: BLE001
    log("smoke: FAILED %r" % (e,))
os.environ["WCTB_PENDING_HW"] = "1"
import pytest  # noqa: E402

rc = pytest.main(["-q", "-ra", "--maxfail=50", "-m", "gpu and pending_hw", "-p", "no:cacheprovider", os.path.join(ROOT, "tests"),
                  "--junitxml", os.path.join(ROOT, "gpurun_out", "pending_hw.xml")] + sys.argv[1:])
log("pytest pending_hw exit code %s" % rc)
