"""Throughput mode (collaborative_distillation_b200/pipeline.py): a sequence of DIFFERENT host pairs through
`WCT.pipeline()` must give, pair by pair, what one blocking `WCT.stylize` call per pair gives -- the staging slots, the three
streams and the event hand-over must never mix two pairs up, with and without the CUDA graph."""
import os
from types import SimpleNamespace

import pytest
import torch

import collaborative_distillation_b200 as P

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wct():
    P.set_precision("h2")
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    P.weights.load_npz_into(w, os.path.join(ROOT, "tests", "golden", "weights_16x.npz"))
    return w.cuda()


@pytest.mark.parametrize("use_graph", [True, False])
def test_pipeline_equals_blocking_calls(use_graph):
    w = _wct()
    w.use_graph = use_graph
    g = torch.Generator().manual_seed(7)
    pairs = [(torch.rand(1, 3, 208, 272, generator=g).pin_memory(), torch.rand(1, 3, 160, 192, generator=g).pin_memory()) for _ in range(5)]
    want = [w.stylize(c.cuda(), s.cuda()).cpu() for c, s in pairs]
    outs = [torch.empty(1, 3, 208, 272).pin_memory() for _ in pairs]
    pipe = w.pipeline()
    evs = [pipe.submit(c, s, o)[1] for (c, s), o in zip(pairs, outs)]
    pipe.drain()
    assert all(e.query() for e in evs)
    for i, (o, r) in enumerate(zip(outs, want)):
        # same kernels on the same data; only the order of the statistics' atomics may differ between two runs
        assert (o - r).abs().max().item() <= 1e-4, i
    # distinct pairs really gave distinct images (a slot mix-up would repeat one)
    assert (outs[0] - outs[1]).abs().max().item() > 1e-2


def test_pipeline_shape_change_and_device_result():
    w = _wct()
    g = torch.Generator().manual_seed(8)
    pipe = w.pipeline(depth=3)
    res = []
    for (H, W) in ((96, 128), (96, 128), (128, 160), (96, 128)):
        c, s = torch.rand(1, 3, H, W, generator=g).pin_memory(), torch.rand(1, 3, 80, 96, generator=g).pin_memory()
        img, ev = pipe.submit(c, s, None)                  # no host buffer: the device image is the result
        res.append((c, s, img, ev))
    pipe.drain()
    for c, s, img, ev in res:
        assert ev.query()
        assert (img.cpu() - w.stylize(c.cuda(), s.cuda()).cpu()).abs().max().item() <= 1e-4
