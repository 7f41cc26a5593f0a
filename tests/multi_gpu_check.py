#!/usr/bin/env python
"""Multi-GPU tile-invariance check (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py [--native-halo]
The strip-sharded stylization must equal the single-GPU result (same kernels per pixel; statistics differ only by
fp64 summation order)."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import collaborative_distillation_b200 as P  # noqa: E402
from collaborative_distillation_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for precision, tol in (("fp32", 2e-4), ("tf32", 5e-3)):
        P.set_precision(precision)
        wct = P.WCT(SimpleNamespace(mode="16x", numpy=False))
        P.weights.load_npz_into(wct, os.path.join(ROOT, "tests", "golden", "weights_16x.npz"))
        wct = wct.to(dev)
        g = torch.Generator().manual_seed(0)
        Wc = 336 * world + 8          # not a multiple of 16: exercises the floor-pool remainder on the last strip
        content = torch.rand(1, 3, 200, Wc, generator=g)
        style = torch.rand(1, 3, 180, 352 * world, generator=g)
        ref = None
        if rank == 0:
            wct.dist = None
            wct.fast_stats = False             # the sharded path always uses the fp64 Gram (partition independent)
            wct.eig_early = 1e-4               # ... and the tight eigensolver early stop (WCT._early)
            ref = wct.stylize(content.to(dev), style.to(dev))
        grp = parallel.StripGroup(native_halo="--native-halo" in sys.argv)   # libwctb halo pack/unpack instead of torch slicing
        wct.dist = grp
        own = grp.stylize(wct.style_transfer_stage, "16x",
                          grp.own_slice(content, parallel.strip_cuts(Wc, world), rank).to(dev),
                          grp.own_slice(style, parallel.strip_cuts(style.shape[-1], world), rank).to(dev))
        parts = [None] * world
        dist.all_gather_object(parts, own.cpu())
        if rank == 0:
            got = torch.cat(parts, dim=-1)
            assert got.shape == ref.shape, (got.shape, ref.shape)
            d = (got - ref.cpu()).abs().max().item()
            print("multi_gpu_check[%s]: world=%d shape=%s max|sharded - single| = %.3g (tol %g)" % (
                precision, world, tuple(got.shape), d, tol))
            ok = ok and d <= tol
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
