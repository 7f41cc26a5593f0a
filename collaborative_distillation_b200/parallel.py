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
                st, s_region, s_count = style_args(s)
                if not cuda:
                    style_res[s] = (stage_fn.style_part(s, st, s_region, s_count), None)
                    return
                with torch.cuda.stream(side):
                    if after is not None:
                        side.wait_event(after)
                    res = stage_fn.style_part(s, st, s_region, s_count)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    for t in res:
                        t.record_stream(main)
                    style_res[s] = (res, ev)
            todo = list(dict.fromkeys(stages))
            launch_style(todo[0])
            slots[todo[0]] = todo[1:3]
            for i in range(1, len(todo)):
                if i + 2 < len(todo):
                    slots[todo[i]] = [todo[i + 2]]

        def content_chain(img, Wc_tot):
            for _ in range(num_run):
                for s in stages:
                    h = stage_halo(mode, s)
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
                        out = stage_fn.content_part(s, ext, get_style, alpha, c_region, c_count, before_eig=release)
                    else:
                        st, s_region, s_count = style_args(s)
                        out = stage_fn(s, ext, st, alpha, c_region, s_region, c_count, s_count)
                    Wc_tot = (Wc_tot >> sh) << sh          # floor-pool drops trailing columns of the whole image
                    # floor-pool may have dropped trailing rows/cols (global right/bottom edge only)
                    x1 = min(We - rh, out.shape[-1])
                    img = out[..., lh:x1].contiguous()
            return img

        if split and cuda:
            with torch.cuda.stream(main):
                img = content_chain(img, Wc_tot)
                img.record_stream(cur)
            cur.wait_stream(main)
            cur.wait_stream(side)
            return img
        return content_chain(img, Wc_tot)
