"""Throughput mode of the stylization path for a SEQUENCE of host-resident pairs (the reference's main loop, WCT.py:109-131,
walks a folder of content x style pairs and pays image upload, five stages and image download strictly one after another).

`StylizePipeline` keeps three CUDA streams busy: while pair i runs its five stages on the compute stream, pair i+1 is copied
host -> device on an upload stream and the result of pair i-1 device -> host on a download stream (PCIe is full duplex and the
copy engines are independent of the SMs).  Device staging buffers are double-buffered and guarded by CUDA events only -- the
host thread never blocks until `drain()`.  Every pair is still uploaded, computed and downloaded in full; nothing is cached.

    pipe = wct.pipeline()                              # or parallel.StripGroup.pipeline(wct, mode, Wc, Ws) when sharded
    for content_h, style_h, out_h in pairs:            # pinned host tensors [1,3,H,W] fp32; out_h receives the image
        pipe.submit(content_h, style_h, out_h)
    pipe.drain()
"""
from __future__ import annotations

import torch


class _Slot:
    def __init__(self):
        self.c = self.s = None
        self.uploaded = torch.cuda.Event()
        self.consumed = torch.cuda.Event()      # the compute stream has finished reading c / s
        self.result = None                      # keeps the device image alive until its download has been enqueued


class StylizePipeline:
    def __init__(self, run, depth: int = 2, device=None):
        """run(content_dev, style_dev) -> image_dev on the current stream (WCT.stylize, or a StripGroup.stylize closure)"""
        self.run = run
        self.depth = max(2, int(depth))
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            self.up = torch.cuda.Stream()
            self.down = torch.cuda.Stream()
        self.slots = [_Slot() for _ in range(self.depth)]
        self.n = 0

    def _staging(self, slot, content_h, style_h):
        """(re)allocate the device staging buffers of a slot -> True when a buffer is new (its first use on the upload stream must
        be ordered after whatever the allocating stream did with that memory before)"""
        fresh = False
        if slot.c is None or slot.c.shape != content_h.shape:
            slot.c = torch.empty(content_h.shape, dtype=torch.float32, device=self.device)
            slot.c.record_stream(self.up)
            fresh = True
        if slot.s is None or slot.s.shape != style_h.shape:
            slot.s = torch.empty(style_h.shape, dtype=torch.float32, device=self.device)
            slot.s.record_stream(self.up)
            fresh = True
        return fresh

    @torch.no_grad()
    def submit(self, content_h: torch.Tensor, style_h: torch.Tensor, out_h: torch.Tensor = None):
        """enqueue one pair; returns (image_dev, done_event).  out_h (pinned host tensor at least as large as the result) receives
        the image asynchronously: valid after done_event / drain()."""
        cur = torch.cuda.current_stream(self.device)
        slot = self.slots[self.n % self.depth]
        self.n += 1
        fresh = self._staging(slot, content_h, style_h)
        with torch.cuda.stream(self.up):
            self.up.wait_event(slot.consumed)            # the pair that used this slot `depth` submissions ago has been read
            if fresh:
                self.up.wait_stream(cur)
            slot.c.copy_(content_h, non_blocking=True)
            slot.s.copy_(style_h, non_blocking=True)
            slot.uploaded.record(self.up)
        cur.wait_event(slot.uploaded)
        img = self.run(slot.c, slot.s)
        slot.consumed.record(cur)
        done = torch.cuda.Event()
        if out_h is not None:
            self.down.wait_stream(cur)
            with torch.cuda.stream(self.down):
                out_h[..., :img.shape[-2], :img.shape[-1]].copy_(img, non_blocking=True)
                done.record(self.down)
            img.record_stream(self.down)
        else:
            done.record(cur)
        slot.result = img
        return img, done

    def drain(self):
        """block the host until every submitted pair has been computed and downloaded"""
        self.down.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
        for s in self.slots:
            s.result = None
