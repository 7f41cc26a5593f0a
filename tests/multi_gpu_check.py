#!/usr/bin/env python
"""Multi-GPU tile-invariance check (run under torchrun on N GPUs; tests/test_multi_gpu.py launches it on 2):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
The strip-sharded stylization (halo exchange + statistic all-reduces, overlapped two-stream schedule) against the DEFAULT
single-GPU output of the same engine, and both against the CPU oracle.  Per pixel the convolution arithmetic is identical;
the statistics differ by summation order (fp64 Gram) or by the partition of fp32 partial sums (h2 engine, large maps)."""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import collaborative_distillation_b200 as P  # noqa: E402
from collaborative_distillation_b200 import parallel  # noqa: E402
from oracle import wct_oracle as O  # noqa: E402

# (precision, max |sharded - single GPU default|, rms vs oracle) bounds; measured values in profiles/r02_multi_gpu_check.txt
BOUNDS = {"h2": (2e-3, 3e-4), "fp32": (2e-3, 3e-4)}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    big = "--big" in sys.argv
    ok, report = True, []
    wpath = os.path.join(ROOT, "tests", "golden", "weights_16x.npz")
    grp = parallel.StripGroup(native_halo="--native-halo" in sys.argv)     # one group for the whole run
    grp.peer_halo = "--no-peer" not in sys.argv           # fused tail writes the neighbours' halos over NVLink (default) vs NCCL exchange
    grp.use_graph = "--no-graph" not in sys.argv and os.environ.get("WCTB_SHARD_GRAPH", "1") == "1"   # captured step (h2 engine) unless told otherwise
    for precision in (("h2",) if big else ("h2", "fp32")):
        P.set_precision(precision)
        wct = P.WCT(SimpleNamespace(mode="16x", numpy=False))
        P.weights.load_npz_into(wct, wpath)
        wct = wct.to(dev)
        g = torch.Generator().manual_seed(0)
        # width not a multiple of 16: exercises the floor-pool remainder on the last strip.  --big: maps large enough for the
        # fp32-product Gram of the h2 engine (>= 65536 feature pixels), i.e. the partition-dependent statistics path
        H, Wc, Hs, Ws = (1040, 544 * world + 8, 600, 512 * world) if big else (200, 336 * world + 8, 180, 352 * world)
        content = torch.rand(1, 3, H, Wc, generator=g)
        style = torch.rand(1, 3, Hs, Ws, generator=g)
        ref = oracle = None
        if rank == 0:
            wct.dist = None
            ref = wct.stylize(content.to(dev), style.to(dev)).cpu()          # the default single-GPU path (graph, fast stats, ...)
            oracle = O.stylize(O.load_weights_npz(wpath), "16x", content, style)
        wct.dist = grp
        c_own = grp.own_slice(content, parallel.strip_cuts(Wc, world), rank).to(dev)
        s_own = grp.own_slice(style, parallel.strip_cuts(Ws, world), rank).to(dev)
        # default: the step is captured in a CUDA graph (eager pass, capture, replay); a second replay and the eager schedule
        # must reproduce it (up to the summation order of the statistics' atomics)
        own = grp.stylize(wct, "16x", c_own, s_own, content_width=Wc, style_width=Ws)
        own_r = grp.stylize(wct, "16x", c_own, s_own, content_width=Wc, style_width=Ws)
        own_e = grp.stylize(wct, "16x", c_own, s_own, content_width=Wc, style_width=Ws, use_graph=False)
        graphed = precision == "h2" and any(e[0] is not None for e in grp._graphs.values())
        dg = torch.tensor([(own_r - own).abs().max().item(), (own_e - own).abs().max().item()], dtype=torch.float64, device=dev)
        dist.all_reduce(dg, op=dist.ReduceOp.MAX)
        # the legacy one-call-per-stage executor must agree with the overlapped one
        own2 = grp.stylize(wct.style_transfer_stage, "16x", grp.own_slice(content, parallel.strip_cuts(Wc, world), rank).to(dev),
                           grp.own_slice(style, parallel.strip_cuts(Ws, world), rank).to(dev))
        full = grp.gather_strips(own)
        full2 = grp.gather_strips(own2)
        if rank == 0:
            got = full.cpu()
            assert got.shape == ref.shape, (got.shape, ref.shape)
            d = (got - ref).abs().max().item()
            d2 = (full2.cpu() - got).abs().max().item()
            e = got - oracle
            rms, mx = e.pow(2).mean().sqrt().item(), e.abs().max().item()
            tol_d, tol_rms = BOUNDS[precision]
            dl = (full2.cpu() - ref).abs().max().item()
            line = ("multi_gpu_check[%s%s%s]: world=%d shape=%s max|sharded - single| = %.3g (tol %g)  overlapped vs per-stage executor %.3g  "
                    "(per-stage executor vs single %.3g)  sharded vs oracle rms %.3g max %.3g (rms tol %g)  graph=%s replay vs replay %.3g  "
                    "replay vs eager %.3g"
                    % (precision, " big" if big else "", "" if grp.peer_halo else " no-peer", world, tuple(got.shape), d, tol_d, d2, dl, rms, mx, tol_rms,
                       graphed, dg[0].item(), dg[1].item()))
            print(line, flush=True)
            report.append(line)
            ok = ok and d <= tol_d and rms <= tol_rms and d2 <= tol_d and dg.max().item() <= tol_d
    if rank == 0:
        out = os.environ.get("WCTB_CHECK_OUT")
        if out:
            with open(out, "w") as f:
                json.dump({"ok": ok, "lines": report}, f)
    parallel.shutdown(0 if ok else 1)      # captured steps hold NCCL nodes: leave without destructors (see parallel.shutdown)
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
