"""Tile invariance on real hardware: spawns 2 NCCL ranks (torchrun) of tests/multi_gpu_check.py and checks that the strip-sharded
stylization equals the default single-GPU output (bound stated in multi_gpu_check.BOUNDS) and the CPU oracle, and that the
captured step (one CUDA graph per rank, h2 engine) replays to what the eager schedule computes.
Skipped when fewer than 2 GPUs are visible (the single-GPU round-end run); the host logic is covered on CPU by
tests/test_strip_parallel_gloo.py (gloo, world 2 and 3)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("extra", [["--no-peer"], [], ["--big"]])
def test_sharded_equals_single_gpu_two_ranks(tmp_path, extra):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "check.json")
    env = dict(os.environ, WCTB_CHECK_OUT=out)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")] + extra
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=360)      # a healthy run takes 20-40 s
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    rep = json.load(open(out))
    assert rep["ok"], rep
