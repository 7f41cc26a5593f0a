"""world_size-2 (and 3) gloo tests of the strip-parallel driver on CPU: partition bookkeeping, halo exchange and
statistic all-reduces must reproduce the single-process oracle (tile invariance).  The per-stage executor here is
a CPU restatement of the product's moment-based algorithm built from oracle pieces (tests may use the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from collaborative_distillation_b200 import parallel  # noqa: E402
from oracle import wct_oracle as O  # noqa: E402


def test_halos_and_cuts():
    assert [parallel.stage_halo("16x", s) for s in (5, 4, 3, 2, 1)] == [160, 64, 32, 16, 16]
    assert parallel.stage_halo("original", 5) == 160
    assert parallel.strip_cuts(10240, 8) == [1280 * i for i in range(9)]
    c = parallel.strip_cuts(2000, 3)
    assert c[0] == 0 and c[-1] == 2000 and all(v % 16 == 0 for v in c[1:-1])
    with pytest.raises(ValueError):
        parallel.strip_cuts(20, 4)


def cpu_stage_fn(weights, mode, group):
    def moments(F, region, count):
        y0, y1, x0, x1 = region
        x = F[:, y0:y1, x0:x1].double().reshape(F.shape[0], -1)
        s = x.sum(1)
        if group is not None:
            group.allreduce_(s)
        n = float(count if count is not None else x.shape[1])
        mean = s / n
        xc = x - mean[:, None]
        g = xc @ xc.t()
        if group is not None:
            group.allreduce_(g)
        return n, mean, g

    def fn(stage, content, style, alpha, c_region, s_region, c_count=None, s_count=None):
        sh = stage - 1
        with torch.no_grad():
            cF = O.encoder_forward(weights["e%d" % stage], mode, stage, content).squeeze(0)
            sF = O.encoder_forward(weights["e%d" % stage], mode, stage, style).squeeze(0)
            reg = lambda r: tuple(v >> sh for v in r)
            nc, cm, cg = moments(cF, reg(c_region), c_count)
            ns, sm, sg = moments(sF, reg(s_region), s_count)
            ce, cv = torch.linalg.eigh(cg / (nc - 1))
            se, sv = torch.linalg.eigh(sg / (ns - 1))
            kc, ks = ce > 1e-7 * ce.max(), se > 1e-7 * se.max()
            Wm = (cv[:, kc] * ce[kc].pow(-0.5)) @ cv[:, kc].t()
            Cm = (sv[:, ks] * se[ks].pow(0.5)) @ sv[:, ks].t()
            M = alpha * (Cm @ Wm) + (1 - alpha) * torch.eye(cF.shape[0], dtype=torch.float64)
            b = alpha * sm + (1 - alpha) * cm
            x = cF.double().reshape(cF.shape[0], -1)
            cs = (M @ (x - cm[:, None]) + b[:, None]).float().view_as(cF).unsqueeze(0)
            return O.decoder_forward(weights["d%d" % stage], mode, stage, cs)
    return fn


class CpuExecutor:
    """the same algorithm as cpu_stage_fn in the two halves the overlapped strip driver uses (WCT.style_part / content_part)"""

    def __init__(self, weights, mode, group):
        self.w, self.mode, self.group = weights, mode, group

    def _moments(self, F, region, count):
        y0, y1, x0, x1 = region
        x = F[:, y0:y1, x0:x1].double().reshape(F.shape[0], -1)
        s = x.sum(1)
        self.group.allreduce_(s)
        mean = s / float(count)
        xc = x - mean[:, None]
        g = xc @ xc.t()
        self.group.allreduce_(g)
        return float(count), mean, g

    def style_part(self, stage, style, s_region, s_count):
        with torch.no_grad():
            sF = O.encoder_forward(self.w["e%d" % stage], self.mode, stage, style).squeeze(0)
            ns, sm, sg = self._moments(sF, tuple(v >> (stage - 1) for v in s_region), s_count)
            se, sv = torch.linalg.eigh(sg / (ns - 1))
        return sm, se, sv

    def content_part(self, stage, content, style_res, alpha, c_region, c_count, before_eig=None, tail_shard=None):
        with torch.no_grad():
            cF = O.encoder_forward(self.w["e%d" % stage], self.mode, stage, content).squeeze(0)
            nc, cm, cg = self._moments(cF, tuple(v >> (stage - 1) for v in c_region), c_count)
            if before_eig is not None:
                before_eig()
            ce, cv = torch.linalg.eigh(cg / (nc - 1))
            sm, se, sv = style_res() if callable(style_res) else style_res
            kc, ks = ce > 1e-7 * ce.max(), se > 1e-7 * se.max()
            Wm = (cv[:, kc] * ce[kc].pow(-0.5)) @ cv[:, kc].t()
            Cm = (sv[:, ks] * se[ks].pow(0.5)) @ sv[:, ks].t()
            M = alpha * (Cm @ Wm) + (1 - alpha) * torch.eye(cF.shape[0], dtype=torch.float64)
            b = alpha * sm + (1 - alpha) * cm
            x = cF.double().reshape(cF.shape[0], -1)
            cs = (M @ (x - cm[:, None]) + b[:, None]).float().view_as(cF).unsqueeze(0)
            return O.decoder_forward(self.w["d%d" % stage], self.mode, stage, cs)


def _worker(rank, world, port, content, style, stages, out_path, split=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        weights = O.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_16x.npz"))
        grp = parallel.StripGroup()
        ccuts = parallel.strip_cuts(content.shape[-1], world)
        scuts = parallel.strip_cuts(style.shape[-1], world)
        fn = CpuExecutor(weights, "16x", grp) if split else cpu_stage_fn(weights, "16x", grp)
        kw = dict(content_width=content.shape[-1], style_width=style.shape[-1]) if split else {}    # host-known widths: no all-reduce
        own = grp.stylize(fn, "16x", grp.own_slice(content, ccuts, rank), grp.own_slice(style, scuts, rank), alpha=1.0,
                          stages=stages, **kw)
        parts = [None] * world
        dist.all_gather_object(parts, own.numpy())
        if rank == 0:
            np.save(out_path, np.concatenate(parts, axis=-1))
        # CPU strips never take the captured-step path (there is nothing to capture), whatever the group's default says
        assert grp.use_graph in (True, False) and not grp._graphs and not parallel._KEEP_ALIVE
    finally:
        parallel.shutdown()          # no captured steps -> an ordinary destroy_process_group()
        assert not dist.is_initialized()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,Wc,Ws,stages,split", [(2, 704, 672, (5, 4), False), (3, 208, 170, (3, 2, 1), False),
                                                      (2, 704, 672, (5, 4), True), (3, 208, 170, (3, 2, 1), True)])
def test_strip_parallel_equals_single_process(tmp_path, world, Wc, Ws, stages, split):
    g = torch.Generator().manual_seed(3)
    content = torch.rand(1, 3, 40, Wc, generator=g)
    style = torch.rand(1, 3, 36, Ws, generator=g)
    weights = O.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_16x.npz"))
    torch.set_num_threads(4)
    # single process, moment-based executor == plain oracle (validates the executor itself)
    fn = cpu_stage_fn(weights, "16x", None)
    img = content
    for s in stages:
        img = fn(s, img, style, 1.0, (0, img.shape[-2], 0, img.shape[-1]), (0, style.shape[-2], 0, style.shape[-1]))
    ref = O.stylize(weights, "16x", content, style, stages=stages)
    assert (img - ref).abs().max().item() <= 1e-4
    out_path = str(tmp_path / "out.npy")
    mp.spawn(_worker, args=(world, _free_port(), content, style, stages, out_path, split), nprocs=world, join=True)
    got = torch.from_numpy(np.load(out_path))
    assert got.shape == img.shape
    # same algorithm per pixel; on CPU oneDNN picks width-dependent conv blockings, so strips differ from the
    # full image by fp32 rounding (amplified by the whitening): observed 5e-5 on a [0,1.4] image.
    assert (got - img).abs().max().item() <= 2e-4


@pytest.mark.parametrize("stage", [5, 4, 3, 2, 1])
def test_stage_halo_covers_the_receptive_field(stage):
    """Impulse test on the CPU oracle: perturbing input column x0 must not change decoder(encoder(.)) outputs farther
    than stage_halo(stage) columns away -- the guarantee the strip driver relies on when it crops the halo."""
    weights = O.load_weights_npz(os.path.join(ROOT, "tests", "golden", "weights_16x.npz"))
    halo = parallel.stage_halo("16x", stage)
    W = 2 * halo + 64
    g = torch.Generator().manual_seed(stage)
    x = torch.rand(1, 3, 32, W, generator=g)
    x0 = W // 2
    x2 = x.clone()
    x2[..., x0] += 0.5
    torch.set_num_threads(4)
    with torch.no_grad():
        f = lambda t: O.decoder_forward(weights["d%d" % stage], "16x", stage, O.encoder_forward(weights["e%d" % stage], "16x", stage, t))
        a, b = f(x), f(x2)
    diff = (a - b).abs().amax(dim=(0, 1, 2))           # per output column
    changed = torch.nonzero(diff > 0).flatten()
    assert changed.numel() > 0
    reach = max(x0 - int(changed.min()), int(changed.max()) - x0)
    assert reach <= halo, "stage %d: influence reaches %d columns, halo is %d" % (stage, reach, halo)
    assert reach > halo - 32 or stage <= 2, "halo %d is far larger than the measured reach %d" % (halo, reach)
