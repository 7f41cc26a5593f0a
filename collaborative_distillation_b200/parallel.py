"""Multi-GPU sharding of the stylization path: spatial strips + halos, one process per GPU.

The reference is single-GPU.  Here an ultra-resolution content image (and the style image) is cut
along W into `world` contiguous strips whose cuts are multiples of 16 px (2^4 = total pooling factor), so no
pool window / upsample pair straddles a seam and the floor-pool shape chain equals the single-GPU one.

Per stage k (WCT.py:98-106) every rank
  1. receives a `halo(k)`-px image halo from its neighbours (one send/recv pair per side, NCCL over NVLink),
     halo(k) = encoder + decoder receptive field of stage k rounded up to 16 (160/64/32/16/16 px for k=5..1),
  2. runs encoder -> statistics -> eigensolve -> apply -> decoder on the extended strip; reflection padding
     happens only at true image borders, the seam-side error stays inside the halo and is cropped away,
  3. all-reduces the statistics (sum x: C doubles, centred Gram: C*C doubles) over its OWN strip only,
so the result is tile-invariant: identical convolution arithmetic per pixel, statistics equal up to fp64
summation order.  No feature map is ever gathered.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import arch


def stage_halo(mode: str, stage: int) -> int:
    """image-space halo (px, multiple of 16) that makes the own strip of stage `stage` exact."""
    h = 0
    for L in reversed(arch.decoder_layers(mode, stage)):
        if L["up_after"]:
            h = (h + 1) // 2
        h += 1
    for L in reversed(arch.encoder_layers(mode, stage)):
        if L["pool_after"]:
            h *= 2
        h += 1
    return ((h + 15) // 16) * 16


def strip_cuts(W: int, world: int):
    """cut positions [c0=0, c1, ..., c_world=W]; interior cuts are multiples of 16."""
    cuts = [0]
    for i in range(1, world):
        cuts.append(int(round(i * W / world / 16.0)) * 16)
    cuts.append(W)
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b <= a:
            raise ValueError("image width %d too small for %d strips" % (W, world))
    return cuts


_KEEP_ALIVE = []   # captured sharded steps (CUDA graphs with NCCL nodes) stay alive until the process exits


def shutdown(code: int = 0):
    """End a multi-GPU process that captured sharded steps.  Measured on 2 B200s (profiles/r02_multi_gpu_graph.txt): with CUDA
    graphs holding NCCL nodes alive, `dist.destroy_process_group()` / interpreter teardown does not return (the communicator
    waits for the graphs, the graphs for the communicator).  So: synchronize, agree that every rank is done, flush, and leave
    without running destructors.  Without captured steps this is an ordinary destroy_process_group()."""
    import sys
    if dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if not _KEEP_ALIVE:
            dist.destroy_process_group()
            return
    elif not _KEEP_ALIVE:
        return
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


class PeerHalo:
    """Peer-mapped next-stage strip buffers for the fused tail kernel (compute + halo exchange in one kernel).

    Every rank owns two device buffers (double buffer: a neighbour may already write stage k-1's halo while the owner still
    reads stage k's strip) and opens its neighbours' buffers through CUDA IPC, so `wctb_conv_tail_h2_sharded` can store the
    seam-side columns of its output straight into the neighbours' next-stage strips over NVLink.  One tiny all-reduce per stage
    is the cross-rank barrier (stream-ordered: every rank's tail kernel has completed before any rank's next encoder starts)."""

    def __init__(self, group, nbytes: int, device):
        import torch.distributed as dist
        self.group, self.rank, self.world = group, dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = int(nbytes)
        self.local = [torch.empty(self.nbytes, dtype=torch.uint8, device=device) for _ in range(2)]
        mine = [t.untyped_storage()._share_cuda_() for t in self.local]
        table = [None] * self.world
        dist.all_gather_object(table, mine, group=group)
        self.peer = {}
        self._keep = []
        me = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        for r in (self.rank - 1, self.rank + 1):
            if 0 <= r < self.world:
                if table[r][0][0] != me and not torch.cuda.can_device_access_peer(me, table[r][0][0]):
                    raise RuntimeError("no peer access from cuda:%d to cuda:%d" % (me, table[r][0][0]))
                bufs = []
                for h in table[r]:
                    # open the neighbour's allocation in THIS device's context (cudaIpcOpenMemHandle with
                    # cudaIpcMemLazyEnablePeerAccess): the mapping is then addressable by kernels running on this GPU
                    with torch.cuda.device(me):
                        st = torch.UntypedStorage._new_shared_cuda(me, *h[1:])
                    t = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st)
                    bufs.append(t)
                    self._keep.append(st)
                self.peer[r] = bufs
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def local_view(self, b: int, H: int, We: int) -> torch.Tensor:
        return self.local[b][:3 * H * We * 4].view(torch.float32).view(1, 3, H, We)

    def peer_ptr(self, r: int, b: int) -> int:
        return self.peer[r][b].data_ptr()

    def barrier(self):
        import torch.distributed as dist
        dist.all_reduce(self.flag, group=self.group)


class StripGroup:
    """Strip-parallel driver.  `stage_fn(stage, content_ext, style_ext, alpha, c_region, s_region, c_count, s_count) -> image_ext`
    is `WCT.style_transfer_stage` in production (with `wct.dist = self`), or a CPU restatement in the gloo tests."""

    def __init__(self, group=None, native_halo=False):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        # native_halo: pack the send buffers and assemble the extended strip with libwctb's halo kernels
        # (wctb_halo_pack / wctb_halo_unpack) instead of torch slicing + cat.  Same bytes either way; off until its first
        # multi-GPU run is on record (the kernels themselves are covered by the single-GPU tests).
        self.native_halo = native_halo
        # peer_halo: let the fused tail kernel write the next stage's halo straight into the neighbours' buffers (PeerHalo);
        # WCTB_PEER_HALO=0 falls back to pack / send / recv / cat for every stage
        self.peer_halo = os.environ.get("WCTB_PEER_HALO", "1") == "1"
        self._peer = None
        self.counters = {"peer_halo": 0, "nccl_halo": 0}     # halo exchanges by mechanism (bench.py reports them)
        # use_graph: capture the whole sharded step of one input shape (two streams, ~1500 launches, the NCCL statistic
        # all-reduces, the first halo exchange and the peer-halo barriers) in ONE CUDA graph per rank and replay it.  With 8
        # ranks on one host the eager schedule is bound by the host (Python launches at ~8 us each, 8 processes sharing the
        # cores); a replay needs one launch.  Every rank captures the same program, so the NCCL order is identical everywhere.
        self.use_graph = os.environ.get("WCTB_SHARD_GRAPH", "1") == "1"   # processes that captured a step must end with parallel.shutdown()
        self._graphs = {}        # never evicted: a captured step holds NCCL nodes, and its destruction is left to process exit

    # ---- collectives used by WCT._moments
    def allreduce_(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def total_width(self, own_w: int, device) -> int:
        t = torch.tensor([own_w], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    # ---- halo exchange
    def exchange(self, own: torch.Tensor, halo: int):
        """own [1,3,H,w] -> extended strip [1,3,H,lh+w+rh] and (lh, rh) actually attached (0 at true borders)."""
        if self.world == 1 or halo == 0:
            return own, 0, 0
        w = own.shape[-1]
        if w < halo:
            raise ValueError("strip width %d < halo %d: use fewer GPUs for this image" % (w, halo))
        r, n = self.rank, self.world
        native = self.native_halo and own.is_cuda
        if native:
            from . import ops as K
            own = own.contiguous()
            pack = lambda x0: K.halo_pack(own, x0, halo)
        else:
            pack = lambda x0: own[..., x0:x0 + halo].contiguous()
        ops, left, right = [], None, None
        if r > 0:
            left = torch.empty(own.shape[:-1] + (halo,), dtype=own.dtype, device=own.device)
            ops += [dist.P2POp(dist.isend, pack(0), r - 1, self.group),
                    dist.P2POp(dist.irecv, left, r - 1, self.group)]
        if r < n - 1:
            right = torch.empty(own.shape[:-1] + (halo,), dtype=own.dtype, device=own.device)
            ops += [dist.P2POp(dist.isend, pack(w - halo), r + 1, self.group),
                    dist.P2POp(dist.irecv, right, r + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        lh, rh = (halo if left is not None else 0), (halo if right is not None else 0)
        self.counters["nccl_halo"] += 1
        if native:
            ext = torch.empty(own.shape[:-1] + (lh + w + rh,), dtype=own.dtype, device=own.device)
            if left is not None:
                K.halo_unpack(left, ext, 0)
            K.halo_unpack(own, ext, lh)
            if right is not None:
                K.halo_unpack(right, ext, lh + w)
            return ext, lh, rh
        parts = [p for p in (left, own, right) if p is not None]
        return torch.cat(parts, dim=-1).contiguous(), lh, rh

    def gather_strips(self, own: torch.Tensor, dst: int = 0):
        """concatenate every rank's strip [1,3,H,w_r] along W on rank `dst` (widths may differ); other ranks get None.
        Only used at the very end of a stylization (saving the image): features are never gathered."""
        w = torch.tensor([own.shape[-1]], dtype=torch.int64, device=own.device)
        ws = [torch.zeros_like(w) for _ in range(self.world)]
        dist.all_gather(ws, w, group=self.group)
        ws = [int(v.item()) for v in ws]
        wmax = max(ws)
        pad = own.new_zeros(own.shape[:-1] + (wmax,))
        pad[..., :own.shape[-1]] = own
        bufs = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        return torch.cat([b[..., :n] for b, n in zip(bufs, ws)], dim=-1)

    @staticmethod
    def own_slice(full: torch.Tensor, cuts, rank: int):
        return full[..., cuts[rank]:cuts[rank + 1]].contiguous()

    def stylize(self, stage_fn, mode: str, content_own: torch.Tensor, style_own: torch.Tensor, alpha: float = 1.0,
                stages=(5, 4, 3, 2, 1), num_run: int = 1, content_width: int = None, style_width: int = None,
                use_graph: bool = None) -> torch.Tensor:
        """Sharded stylization; see `_stylize_eager` for the arguments.  With `use_graph` (default: self.use_graph) and a `WCT`
        executor on CUDA strips whose whole-image widths are given, the step is captured once per input shape in a CUDA graph
        (after one eager pass that packs weights and opens the peer buffers) and replayed afterwards."""
        use_graph = self.use_graph if use_graph is None else use_graph
        if (use_graph and content_own.is_cuda and style_own.is_cuda and content_width is not None and style_width is not None
                and hasattr(stage_fn, "style_part") and hasattr(stage_fn, "content_part") and getattr(stage_fn, "timeline", None) is None
                and self._graph_engine_ok()):
            return self._stylize_graph(stage_fn, mode, content_own, style_own, alpha, tuple(stages), num_run, content_width, style_width)
        return self._stylize_eager(stage_fn, mode, content_own, style_own, alpha, stages, num_run, content_width, style_width)

    def pipeline(self, wct, mode: str, content_width: int, style_width: int, alpha: float = 1.0, stages=(5, 4, 3, 2, 1),
                 num_run: int = 1, depth: int = 2):
        """per-rank `pipeline.StylizePipeline` around the sharded step: each rank uploads its strips of pair i+1 and downloads its
        strip of result i-1 while pair i is being computed"""
        from .pipeline import StylizePipeline
        return StylizePipeline(lambda c, s: self.stylize(wct, mode, c, s, alpha=alpha, stages=stages, num_run=num_run,
                                                         content_width=content_width, style_width=style_width), depth=depth)

    @staticmethod
    def _graph_engine_ok():
        """the capture is exercised (tests/multi_gpu_check.py, bench.py) with the default h2 engine only"""
        from . import nets
        return nets.get_precision() == "h2"

    def _stylize_graph(self, wct, mode, content_own, style_own, alpha, stages, num_run, content_width, style_width):
        from . import nets, ops
        fp = wct._weights_fingerprint(stages) if hasattr(wct, "_weights_fingerprint") else 0
        key = (mode, tuple(content_own.shape), tuple(style_own.shape), float(alpha), stages, int(num_run), int(content_width), int(style_width),
               nets.get_precision(), bool(self.peer_halo), bool(self.native_halo), torch.cuda.current_device(), fp,
               bool(getattr(wct, "fold_into_decoder", True)), float(getattr(wct, "tau", 0.0)))
        ent = self._graphs.get(key)
        if ent is None:
            sc = torch.empty_like(content_own, memory_format=torch.contiguous_format)
            ss = torch.empty_like(style_own, memory_format=torch.contiguous_format)
            sc.copy_(content_own)
            ss.copy_(style_own)
            args = (wct, mode, sc, ss, alpha, stages, num_run, content_width, style_width)
            self._stylize_eager(*args)               # eager pass: packs weights, sets kernel attributes, opens the peer buffers
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            graph, out, nl, cnt, ok = torch.cuda.CUDAGraph(), None, 0, {}, 1
            c0 = dict(self.counters)
            try:
                n0 = ops.launches()
                # thread_local: the NCCL watchdog thread polls its events while this thread captures
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = self._stylize_eager(*args)
                nl = ops.launches() - n0
                cnt = {k: self.counters[k] - c0[k] for k in c0}
            except Exception as e:  # noqa: BLE001
                print("wct-b200: CUDA graph capture of the sharded step failed on rank %d (%s); running eagerly" % (self.rank, e))
                ok = 0
            torch.cuda.synchronize()
            t = torch.tensor([ok], dtype=torch.int32, device=content_own.device)      # every rank replays, or none does
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            if int(t.item()) == 0:
                graph = None
            ent = (graph, sc, ss, out, nl, cnt)
            self._graphs[key] = ent
            _KEEP_ALIVE.append(ent)
        graph, sc, ss, out, nl, cnt = ent
        if graph is None:
            return self._stylize_eager(wct, mode, content_own, style_own, alpha, stages, num_run, content_width, style_width)
        sc.copy_(content_own, non_blocking=True)
        ss.copy_(style_own, non_blocking=True)
        graph.replay()
        ops.add_launches(nl)
        for k, v in cnt.items():
            self.counters[k] += v
        return out.clone()

    def _stylize_eager(self, stage_fn, mode: str, content_own: torch.Tensor, style_own: torch.Tensor, alpha: float = 1.0,
                       stages=(5, 4, 3, 2, 1), num_run: int = 1, content_width: int = None, style_width: int = None) -> torch.Tensor:
        """content_own / style_own: this rank's strips [1,3,H,w].  Returns this rank's strip of the stylized image.
        `stage_fn` is either a callable `fn(stage, content_ext, style_ext, alpha, c_region, s_region, c_count, s_count)`
        (one fused stage) or an executor with `style_part(stage, style_ext, s_region, s_count)` and
        `content_part(stage, content_ext, style_res, alpha, c_region, c_count)` (a `WCT`): then the style halves run on a side
        stream and are released into the content branch's eigensolve gaps, as on one GPU.
        content_width / style_width: whole-image widths when the caller knows them (saves two tiny all-reduces + host syncs)."""
        hmax = max(stage_halo(mode, s) for s in stages)
        style_ext, s_lh, s_rh = self.exchange(style_own, hmax)
        img = content_own
        # statistics divide by GLOBAL pixel counts
        Wc_tot = int(content_width) if content_width is not None else self.total_width(content_own.shape[-1], content_own.device)
        Ws_tot = int(style_width) if style_width is not None else self.total_width(style_own.shape[-1], style_own.device)
        Hs = style_own.shape[-2]
        split = hasattr(stage_fn, "style_part") and hasattr(stage_fn, "content_part")

        def style_args(s):
            h = stage_halo(mode, s)
            st = style_ext[..., (s_lh - min(s_lh, h)):style_ext.shape[-1] - (s_rh - min(s_rh, h))]
            sl, sr = min(s_lh, h), min(s_rh, h)
            return st.contiguous(), (0, st.shape[-2], sl, st.shape[-1] - sr), (Hs >> (s - 1)) * (Ws_tot >> (s - 1))

        cuda = content_own.is_cuda
        style_res, slots = {}, {}
        if split:
            if cuda:
                cur = torch.cuda.current_stream()
                if getattr(self, "_side", None) is None:
                    self._main, self._side = torch.cuda.Stream(priority=-1), torch.cuda.Stream(priority=0)
                main, side = self._main, self._side
                main.wait_stream(cur)
                side.wait_stream(cur)

            def launch_style(s, after=None):
                if s in style_res:
                    return
                if not cuda:
                    st, s_region, s_count = style_args(s)
                    style_res[s] = (stage_fn.style_part(s, st, s_region, s_count), None)
                    return
                with torch.cuda.stream(side):
                    if after is not None:
                        side.wait_event(after)
                    st, s_region, s_count = style_args(s)        # the strip copy must run on the stream that consumes it
                    res = stage_fn.style_part(s, st, s_region, s_count)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    if not torch.cuda.is_current_stream_capturing():
                        for t in res:
                            t.record_stream(main)
                    style_res[s] = (res, ev)
            todo = list(dict.fromkeys(stages))
            launch_style(todo[0])
            slots[todo[0]] = todo[1:3]
            for i in range(1, len(todo)):
                if i + 2 < len(todo):
                    slots[todo[i]] = [todo[i + 2]]

        # per-rank own widths (every rank can compute every rank's geometry: cuts are a pure function of the total width)
        widths = None
        use_peer = split and cuda and self.peer_halo and self.world > 1 and content_width is not None
        if use_peer:
            cuts = strip_cuts(int(content_width), self.world)
            widths = [cuts[i + 1] - cuts[i] for i in range(self.world)]
            if widths[self.rank] != content_own.shape[-1]:
                use_peer = False                      # caller did not use strip_cuts(): geometry of the neighbours unknown
        if use_peer:
            H0 = content_own.shape[-2]
            need = 3 * H0 * (max(widths) + 2 * hmax) * 4
            if self._peer is None or self._peer.nbytes < need:
                try:
                    self._peer = PeerHalo(self.group, need, content_own.device)
                    ok = 1
                except Exception as e:  # noqa: BLE001  (e.g. no NVLink / IPC not permitted): every rank falls back together
                    print("wct-b200: peer-memory halo unavailable (%s); using the NCCL halo exchange" % (e,))
                    self._peer, ok = None, 0
                t = torch.tensor([ok], dtype=torch.int32, device=content_own.device)
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
                if int(t.item()) == 0:
                    self._peer, self.peer_halo, use_peer = None, False, False

        def content_chain(img, Wc_tot):
            nonlocal widths
            r, n = self.rank, self.world
            pending_ext = None                        # (ext, lh, rh) already assembled in a peer buffer by the previous stage's tail
            pbuf = 0
            seq = [s for _ in range(num_run) for s in stages]
            for idx, s in enumerate(seq):
                h = stage_halo(mode, s)
                if pending_ext is not None:
                    ext, lh, rh = pending_ext
                    pending_ext = None
                else:
                    ext, lh, rh = self.exchange(img, h)
                H, We = ext.shape[-2:]
                c_region = (0, H, lh, We - rh)
                sh = s - 1
                c_count = (H >> sh) * (Wc_tot >> sh)
                if split:
                    # the style stages assigned to this slot are issued BEFORE this stage's content work in program order
                    # (NCCL runs collectives in issue order); on the device they wait for the content statistics
                    pending = slots.pop(s, [])

                    def release():          # content statistics enqueued: the eigensolve gap starts here
                        ev_slot = None
                        if cuda:
                            ev_slot = torch.cuda.Event()
                            ev_slot.record(torch.cuda.current_stream())
                        for ss in pending:
                            launch_style(ss, ev_slot)

                    def get_style():
                        res, ev = style_res[s]
                        if ev is not None:
                            torch.cuda.current_stream().wait_event(ev)
                        return res
                    shard = None
                    if use_peer and idx + 1 < len(seq):
                        # geometry of the NEXT stage's extended strips on this rank and on both neighbours
                        hn = stage_halo(mode, seq[idx + 1])
                        wn = list(widths)
                        wn[-1] = (wn[-1] >> sh) << sh                      # floor-pool drops trailing columns (global right edge)
                        Hn = (H >> sh) << sh
                        geo = lambda q: ((hn if q > 0 else 0), wn[q], (hn if q < n - 1 else 0))
                        lhn, wown, rhn = geo(r)
                        if min(wn) >= hn and wown > 0:
                            out = self._peer.local_view(pbuf, Hn, lhn + wown + rhn)
                            shard = {"out": out, "out_x0": lhn, "own_x0": lh, "own_w": wown, "halo": hn, "peer_l": None, "peer_r": None}
                            if r > 0:
                                ll, wl, rl = geo(r - 1)
                                shard["peer_l"] = (self._peer.peer_ptr(r - 1, pbuf), ll + wl + rl, ll + wl)
                            if r < n - 1:
                                lr, wr, rr = geo(r + 1)
                                shard["peer_r"] = (self._peer.peer_ptr(r + 1, pbuf), lr + wr + rr, 0)
                    out = stage_fn.content_part(s, ext, get_style, alpha, c_region, c_count, before_eig=release, tail_shard=shard)
                    if shard is not None and out is shard["out"]:
                        # the fused tail wrote our strip and both neighbours' halos: one stream-ordered barrier, no exchange
                        self._peer.barrier()
                        self.counters["peer_halo"] += 1
                        widths = wn
                        Wc_tot = (Wc_tot >> sh) << sh
                        pending_ext = (out, lhn, rhn)
                        pbuf ^= 1
                        img = None
                        continue
                else:
                    st, s_region, s_count = style_args(s)
                    out = stage_fn(s, ext, st, alpha, c_region, s_region, c_count, s_count)
                if widths is not None:
                    widths[-1] = (widths[-1] >> sh) << sh
                Wc_tot = (Wc_tot >> sh) << sh          # floor-pool drops trailing columns of the whole image
                # floor-pool may have dropped trailing rows/cols (global right/bottom edge only)
                x1 = min(We - rh, out.shape[-1])
                img = out[..., lh:x1].contiguous()
            return img

        if split and cuda:
            with torch.cuda.stream(main):
                img = content_chain(img, Wc_tot)
                if not torch.cuda.is_current_stream_capturing():
                    img.record_stream(cur)
            cur.wait_stream(main)
            cur.wait_stream(side)
            return img
        return content_chain(img, Wc_tot)
