"""`WCT` -- drop-in for PytorchWCT/util_wct.py:30-223, running on hand-written sm_100a kernels.

Same surface: `WCT(args)` reads args.mode / args.e1..e5 / args.d1..d5 / args.numpy; attributes e1..e5, d1..d5;
`transform(cF, sF, csF, alpha)`; `whiten_and_color(cF, sF)`.  Added: `stylize(content, style, alpha)`, the
fused device-resident 5-stage loop of WCT.py:98-106,120-125 (no host round trips, no empty_cache()).

Differences from the reference, all stated in DESIGN.md:
  * the feature transform runs on the GPU (fp64 statistics + fp64 Jacobi eigensolver, fp32 apply) instead of
    CPU fp64; eigen-directions with eigenvalue <= tau*lambda_max (tau=1e-7) are dropped -- the reference's
    EigenValueThre=1e-100 (util_wct.py:25) never triggers, not even on fp64 SVD noise, so the reference whitens EVERY
    direction to unit variance.  For the exact null space (dead ReLU channels) dropping is the same thing; a genuinely live
    direction with relative variance below 1e-7 (std 3e-4 of the dominant one) would be whitened by the reference and is
    dropped here.  On the goldens and the BASELINE inputs the live spectrum ends >= 6e-4*lambda_max, so results agree to
    the tested tolerance; `wct.tau` is a per-instance knob (set it to ~1e-12 with the fp32 engine to follow the reference
    further down the spectrum);
  * wrong mode raises ValueError after printing the reference's message (the reference calls exit(1), :57-59).
"""
from __future__ import annotations

import collections
import os

import torch
import torch.nn as nn

from . import nets, ops
from ._lib import WctbError

EigenValueThre = 1e-100   # kept for reference-compat; see TAU
TAU = 1e-7                # relative eigenvalue threshold used by the GPU path
NumEigenValue = None      # util_wct.py:26 (30 there; its use at :87,:113 is commented out): keep only this many directions
RatEigenValue = None      # util_wct.py:27 (0.25 there; use at :88,:114 commented out): keep int(C * ratio) directions


class WCT(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        mode = args.mode if getattr(args, "mode", None) is not None else "original"   # util_wct.py:35
        if mode not in nets.ENCODERS:
            print("Wrong mode. Please check.")                                          # util_wct.py:57-59
            raise ValueError("wrong mode %r" % (mode,))
        self.mode = mode
        for k in range(5, 0, -1):
            setattr(self, "e%d" % k, nets.ENCODERS[mode][k - 1](getattr(args, "e%d" % k, None)))
            setattr(self, "d%d" % k, nets.DECODERS[mode][k - 1](getattr(args, "d%d" % k, None)))
        self.tau = TAU
        # early-stop |cos| of the eigensolver sweeps (wctb_eigh_jacobi_tol).  None = by conv precision: 1e-4 with the fp32
        # engine (residual <= ~1e-8, whitening error ~1e-9), 1e-2 with the TF32 engine (whitening error ~2e-6 against
        # 1e-3 of TF32 feature noise; two sweeps fewer on the critical path).  WCTB_EIG_EARLY overrides (A/B runs).
        self.eig_early = float(os.environ["WCTB_EIG_EARLY"]) if os.environ.get("WCTB_EIG_EARLY") else None
        # content-side whitening solver: "jacobi" (default: eigendecomposition, one CTA) or "ns" (pivoted Cholesky +
        # Newton-Schulz on a cooperative grid, C <= 128, no eigenvalue-truncation knobs; opt-in until it has run on hardware)
        self.whiten_solver = os.environ.get("WCTB_WHITEN", "jacobi")
        self.num_eig = NumEigenValue   # per-instance knobs; None = keep all directions (the reference's active behaviour)
        self.rat_eig = RatEigenValue
        self.dist = None          # set by parallel.StripGroup for multi-GPU runs
        self.fold_into_decoder = True   # csF = M(cF - mu) + b folded exactly into the decoder's first conv (no apply pass)
        self.overlap_style = True  # single-GPU stylize(): run the (content-independent) style branch on a side stream
        self.stagger_style = os.environ.get("WCTB_STAGGER", "1") == "1"   # release style stages into the content eigensolve gaps
        self._side = None
        self._main = None
        self._stream_dev = None
        self.fast_stats = True     # TF32 mode, single GPU: fp32-product Gram (see _moments)
        # h2 engine: fp32-product Gram (1e-8 relative, far below the 1e-6 noise of the features themselves) on large maps only;
        # small or nearly rank-deficient maps (HW < 64 C, HW < 65536) keep the fp64 Gram: there its cost is nil and the
        # noise eigenvalues of the fp32 products (1e-8 lambda_max) would sit next to the rank threshold tau
        self.fast_stats_h2 = os.environ.get("WCTB_FAST_STATS_H2", "1") == "1"
        self.use_graph = True      # single-GPU stylize(): capture the two-stream schedule in a CUDA graph per input shape
        self.max_graphs = 4        # captured graphs kept (LRU): each one pins the activations of its input shape in HBM
        self._graphs = collections.OrderedDict()

    def _early(self):
        if self.eig_early is not None:
            return float(self.eig_early)
        # sharded runs keep the tight setting: with 1e-2 a last-bit difference in the all-reduced statistics could flip the
        # sweep count and move the whitening matrix by ~1e-5 between partitions; at 1e-4 such a flip is worth <= 1e-8
        return 1e-2 if (nets.get_precision() == "tf32" and self.dist is None) else 1e-4

    def _use_ns(self, C):
        return self.whiten_solver == "ns" and C <= 128 and self._keep(C) == 0

    def _keep(self, C):
        """number of eigen-directions kept for content and style (0 = all): k = NumEigenValue, or int(C * RatEigenValue)"""
        if self.num_eig is not None:
            return int(self.num_eig)
        if self.rat_eig is not None:
            return int(C * self.rat_eig)
        return 0

    # ------------------------------------------------------------------ statistics -> (M, b, mean_c)
    def _moments(self, x_p4, region, count, gram_out):
        """mean fp64 [C] and centred Gram (into gram_out [C,C] fp64, zeroed by the caller) over the region;
        all-reduced over ranks when sharded.  `count` = number of feature pixels of the WHOLE image (host value)."""
        s = ops.channel_sum(x_p4, region)
        if self.dist is not None:
            self.dist.allreduce_(s)
        mean = s / count
        # fp32-product Gram only on a single GPU: its partial sums depend on the pixel partition (1e-7 relative), which the
        # TF32 pipeline amplifies through the whitening; the sharded path keeps the fp64 Gram so that strips stay
        # tile-invariant (sharded == single GPU to fp64 summation order).
        prec = nets.get_precision()
        C = x_p4.shape[0] * 4
        # TF32 engine: fp32-product Gram only on a single GPU (its 1e-7 partition-dependent differences are amplified by the
        # lossy pipeline: 0.185 max between partitions on a noise image).  h2 engine: large maps on one or many GPUs (the
        # partition dependence, 1e-8 of the Gram, stays below the features' own 1e-6; tests/test_multi_gpu.py states the bound)
        fast = ((prec == "tf32" and self.fast_stats and self.dist is None) or
                (prec == "h2" and self.fast_stats_h2 and count >= 65536 and count >= 64 * C))
        ops.centered_gram(x_p4, mean, region, out=gram_out, fast=fast)
        return mean

    def _wct_params(self, c_p4, s_p4, alpha, c_region=None, s_region=None, c_count=None, s_count=None):
        C = c_p4.shape[0] * 4
        full = lambda t: (0, t.shape[1], 0, t.shape[2])
        c_region = c_region or full(c_p4)
        s_region = s_region or full(s_p4)
        nc = float(c_count if c_count is not None else (c_region[1] - c_region[0]) * (c_region[3] - c_region[2]))
        ns = float(s_count if s_count is not None else (s_region[1] - s_region[0]) * (s_region[3] - s_region[2]))
        if nc < 2 or ns < 2:
            raise WctbError("the covariance divides by HW - 1 (util_wct.py:70,96): feature maps need at least 2 pixels")
        grams = torch.zeros(2, C, C, device=c_p4.device, dtype=torch.float64)
        c_mean = self._moments(c_p4, c_region, nc, grams[0])
        s_mean = self._moments(s_p4, s_region, ns, grams[1])
        if self.dist is not None:
            self.dist.allreduce_(grams)
        scale = [1.0 / (nc - 1.0), 1.0 / (ns - 1.0)]                                   # util_wct.py:70,96
        if self._use_ns(C):
            w_c = ops.whiten_ns(grams[0], scale[0], add_identity=bool(getattr(self.args, "numpy", False)))
            se, sv = ops.eigh_jacobi(grams[1:2], scale[1:2], early_stop=self._early())
            return ops.wct_matrix_w(w_c, c_mean, se[0], sv[0], s_mean, self.tau, alpha)
        if getattr(self.args, "numpy", False):                                          # +I on the content covariance only (util_wct.py:143)
            ce, cv = ops.eigh_jacobi(grams[0:1], scale[0:1], add_identity=True, early_stop=self._early())
            se, sv = ops.eigh_jacobi(grams[1:2], scale[1:2], add_identity=False, early_stop=self._early())
            evals, evecs = torch.cat([ce, se]), torch.cat([cv, sv])
        else:
            evals, evecs = ops.eigh_jacobi(grams, scale, early_stop=self._early())
        k = self._keep(C)
        return ops.wct_matrix(evals[0], evecs[0], c_mean, evals[1], evecs[1], s_mean, self.tau, alpha, k, k)

    # ------------------------------------------------------------------ reference API
    def whiten_and_color(self, cF, sF):
        """cF [C,HWc], sF [C,HWs] -> [C,HWc] (fp64 like the reference, util_wct.py:62-131, 204-208)"""
        dev = cF.device
        c = cF.detach().to("cuda", torch.float32).contiguous()
        s = sF.detach().to("cuda", torch.float32).contiguous()
        C = c.shape[0]
        c4 = ops.nchw_to_p4(c.view(C, 1, -1))
        s4 = ops.nchw_to_p4(s.view(C, 1, -1))
        m, b, mc = self._wct_params(c4, s4, 1.0)
        out = ops.p4_to_nchw(ops.wct_apply(c4, m, b, mc)).view(C, -1)
        return out.double().to(dev)

    def transform(self, cF, sF, csF, alpha):
        """cF [C,H,W], sF [C,H1,W1] (CPU or CUDA) -> fills and returns the caller's csF as [1,C,H,W] fp32
        (util_wct.py:210-223)."""
        c = cF.detach().to("cuda", torch.float32).contiguous()
        s = sF.detach().to("cuda", torch.float32).contiguous()
        c4, s4 = ops.nchw_to_p4(c), ops.nchw_to_p4(s)
        m, b, mc = self._wct_params(c4, s4, float(alpha))
        out = ops.p4_to_nchw(ops.wct_apply(c4, m, b, mc))
        csF.resize_(out.shape).copy_(out)
        return csF

    # ------------------------------------------------------------------ fused device-resident path
    @torch.no_grad()
    def style_transfer_stage(self, stage, content, style, alpha=1.0, c_region=None, s_region=None, c_count=None,
                             s_count=None):
        """One styleTransfer(wct.eK, wct.dK, cImg, sImg, csF) of WCT.py:98-106, entirely on the device.
        content/style [1,3,H,W] CUDA; when sharded, regions are this rank's own strip (y0,y1,x0,x1) in image
        pixels and counts are the WHOLE image's number of feature pixels at this stage."""
        enc, dec = getattr(self, "e%d" % stage), getattr(self, "d%d" % stage)
        sh = stage - 1
        s4 = enc.forward_p4(style)
        c4, c8 = self._encode_content(enc, dec, content)
        reg = lambda r: None if r is None else tuple(v >> sh for v in r)
        m, b, mc = self._wct_params(c4, s4, float(alpha), reg(c_region), reg(s_region), c_count, s_count)
        del s4
        if self.fold_into_decoder:
            L0 = getattr(dec, dec.layers[0]["name"])
            w, bb = ops.fold_wct_into_conv(L0.weight.detach().contiguous(), L0.bias.detach().contiguous(), m, b, mc)
            return dec.forward_p4(c4 if c8 is None else c8, first_override=(w, bb))
        cs4 = ops.wct_apply(c4, m, b, mc, round_tf32=dec.first_layer_needs_tf32_input())
        del c4
        return dec.forward_p4(cs4)

    # ---- the same stage in two halves, for the strip driver (parallel.StripGroup): the style half does not depend on the
    # content image, so the driver runs it on a side stream / releases it into the content eigensolve gaps
    @torch.no_grad()
    def style_part(self, stage, style, s_region=None, s_count=None):
        """-> (mean fp64 [C], eigenvalues [C], eigenvectors [C,C]) of the style features of this stage; statistics over
        `s_region` (image pixels) only and all-reduced over the ranks when sharded (util_wct.py:93-100)"""
        enc = getattr(self, "e%d" % stage)
        sh = stage - 1
        s4 = enc.forward_p4(style)
        C = s4.shape[0] * 4
        reg = (0, s4.shape[1], 0, s4.shape[2]) if s_region is None else tuple(v >> sh for v in s_region)
        n = float(s_count if s_count is not None else (reg[1] - reg[0]) * (reg[3] - reg[2]))
        gram = torch.zeros(1, C, C, device=s4.device, dtype=torch.float64)
        mean = self._moments(s4, reg, n, gram[0])
        if self.dist is not None:
            self.dist.allreduce_(gram)
        evals, evecs = ops.eigh_jacobi(gram, [1.0 / (n - 1.0)], early_stop=self._early())
        return mean, evals[0], evecs[0]

    @torch.no_grad()
    def content_part(self, stage, content, style_res, alpha=1.0, c_region=None, c_count=None, before_eig=None, tail_shard=None):
        """content half of styleTransfer (WCT.py:98-106) given the style half's result -> stylized image (extended strip).
        style_res may be a callable returning it (evaluated after the eigensolve is enqueued); before_eig() is called when the
        content statistics are enqueued, i.e. where the single-CTA eigensolve starts and the GPU has room for other work.
        tail_shard: output placement for the fused tail kernel incl. the neighbours' peer pointers (ops.conv_tail_h2); the
        returned tensor IS tail_shard["out"] when the decoder used it."""
        enc, dec = getattr(self, "e%d" % stage), getattr(self, "d%d" % stage)
        sh = stage - 1
        c4, c8 = self._encode_content(enc, dec, content)
        C = c4.shape[0] * 4
        reg = (0, c4.shape[1], 0, c4.shape[2]) if c_region is None else tuple(v >> sh for v in c_region)
        n = float(c_count if c_count is not None else (reg[1] - reg[0]) * (reg[3] - reg[2]))
        gram = torch.zeros(1, C, C, device=c4.device, dtype=torch.float64)
        c_mean = self._moments(c4, reg, n, gram[0])
        if self.dist is not None:
            self.dist.allreduce_(gram)
        numpy_variant = bool(getattr(self.args, "numpy", False))
        if before_eig is not None:
            before_eig()
        c_e, c_v = ops.eigh_jacobi(gram, [1.0 / (n - 1.0)], add_identity=numpy_variant, early_stop=self._early())
        s_mean, s_e, s_v = style_res() if callable(style_res) else style_res
        k = self._keep(C)
        m, b, mc = ops.wct_matrix(c_e[0], c_v[0], c_mean, s_e, s_v, s_mean, self.tau, float(alpha), k, k)
        if self.fold_into_decoder:
            L0 = getattr(dec, dec.layers[0]["name"])
            w, bb = ops.fold_wct_into_conv(L0.weight.detach().contiguous(), L0.bias.detach().contiguous(), m, b, mc)
            return dec.forward_p4(c4 if c8 is None else c8, first_override=(w, bb), tail_shard=tail_shard)
        cs4 = ops.wct_apply(c4, m, b, mc, round_tf32=dec.first_layer_needs_tf32_input())
        del c4
        return dec.forward_p4(cs4, tail_shard=tail_shard)

    def _encode_content(self, enc, dec, img):
        """content features as (fp32 P4 for the statistics, H8 for the decoder's folded first conv | None)"""
        if nets.get_precision() == "h2":
            return enc.forward_feat(img, want_h8=self.fold_into_decoder)
        return enc.forward_p4(img, round_output=self.fold_into_decoder and dec.first_layer_needs_tf32_input()), None

    # ---- two-stream schedule: the style branch (encoder, statistics, eigensolve of every stage) does not depend on the
    # content image, so it runs on a side stream and overlaps the content branch -- in particular the single-CTA
    # eigensolves, which otherwise leave 147 SMs idle.  Results are handed over with CUDA events.
    def _eig_one(self, x_p4, add_identity=False):
        C = x_p4.shape[0] * 4
        n = float(x_p4.shape[1] * x_p4.shape[2])
        gram = torch.zeros(1, C, C, device=x_p4.device, dtype=torch.float64)
        mean = self._moments(x_p4, (0, x_p4.shape[1], 0, x_p4.shape[2]), n, gram[0])
        evals, evecs = ops.eigh_jacobi(gram, [1.0 / (n - 1.0)], add_identity=add_identity, early_stop=self._early())
        return mean, evals[0], evecs[0]

    @torch.no_grad()
    def prepare_style(self, style, stages=(5, 4, 3, 2, 1)):
        """The content-independent half of the path for one style image: per stage (mean, eigenvalues, eigenvectors) of the
        style features (util_wct.py:93-100).  Pass the result as `style_cache=` to stylize() to reuse it across content
        images (WCT.py pairs every content with every style and re-encodes the style each time)."""
        style = style.to("cuda", torch.float32)
        out = {}
        for s in stages:
            out[s] = self._eig_one(getattr(self, "e%d" % s).forward_p4(style))
        return out

    def _mark(self, stage, name):
        """optional CUDA-event timeline of the critical path (tools/stage_timeline.py sets self.timeline = [])"""
        tl = getattr(self, "timeline", None)
        if tl is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            tl.append((stage, name, e))

    @torch.no_grad()
    def _stylize_two_streams(self, content, style, alpha, num_run, stages, style_cache=None):
        """Content branch on a HIGH-priority stream, the whole (content-independent) style branch -- encoder,
        statistics and eigensolve of every stage -- on a low-priority stream: style work fills the SMs whenever the
        critical path leaves them idle (notably during the single-CTA content eigensolves) without delaying it."""
        cur = torch.cuda.current_stream()
        dev = torch.cuda.current_device()
        if self._side is None or self._stream_dev != dev:        # streams belong to a device
            self._main = torch.cuda.Stream(priority=-1)
            self._side = torch.cuda.Stream(priority=0)
            self._stream_dev = dev
        main, side = self._main, self._side
        main.wait_stream(cur)
        side.wait_stream(cur)
        style_res = {}
        if style_cache is not None:
            style_res = {s: (style_cache[s], None) for s in stages}
        img0 = None
        if not content.is_cuda:
            # host buffers: the content image goes up first at full PCIe bandwidth (it gates the critical path); the
            # style copy follows on the side stream and overlaps the first content kernels
            with torch.cuda.stream(main):
                img0 = content.to("cuda", torch.float32, non_blocking=True)
                ev_c = torch.cuda.Event()
                ev_c.record(main)
            side.wait_event(ev_c)
        with torch.cuda.stream(side):
            if style_cache is not None:
                pass
            elif not style.is_cuda:
                style = style.to("cuda", torch.float32, non_blocking=True)
        def launch_style(s, after=None):
            """style branch of stage s on the side stream (optionally not before event `after` of the main stream)"""
            with torch.cuda.stream(side):
                if after is not None:
                    side.wait_event(after)
                s4 = getattr(self, "e%d" % s).forward_p4(style)
                res = self._eig_one(s4)
                del s4
                ev = torch.cuda.Event()
                ev.record(side)
                if not torch.cuda.is_current_stream_capturing():
                    for t in res:
                        t.record_stream(main)
                style_res[s] = (res, ev)

        # Staggered schedule: the content eigensolves are single-CTA kernels on the critical path (2.3 ms of a cfg3 step) during
        # which 147 SMs idle, and the persistent conv kernels of the two branches cannot share SMs anyway.  So only the first
        # stage's style work starts at once; the style work of stage S[i+2] (and S[1], S[2] for i = 0) is released when the
        # content branch reaches the eigensolve of stage S[i], i.e. exactly when the GPU would otherwise go idle.
        todo = [s for s in stages if s not in style_res]
        slots = {}
        if style_cache is None and todo:
            launch_style(todo[0])
            if self.stagger_style:
                slots[todo[0]] = todo[1:3]
                for i in range(1, len(todo)):
                    if i + 2 < len(todo):
                        slots[todo[i]] = [todo[i + 2]]
            else:
                for s in todo[1:]:
                    launch_style(s)
        numpy_variant = bool(getattr(self.args, "numpy", False))
        with torch.cuda.stream(main):
            img = content if img0 is None else img0
            for run in range(num_run):
                for s in stages:
                    enc, dec = getattr(self, "e%d" % s), getattr(self, "d%d" % s)
                    mark = self._mark
                    mark(s, "start")
                    c4, c8 = self._encode_content(enc, dec, img)
                    mark(s, "enc")
                    C = c4.shape[0] * 4
                    n = float(c4.shape[1] * c4.shape[2])
                    gram = torch.zeros(1, C, C, device=c4.device, dtype=torch.float64)
                    c_mean = self._moments(c4, (0, c4.shape[1], 0, c4.shape[2]), n, gram[0])
                    mark(s, "stats")
                    if slots.get(s):
                        ev_slot = torch.cuda.Event()
                        ev_slot.record(main)
                        for ss in slots.pop(s):
                            launch_style(ss, ev_slot)
                    use_ns = self._use_ns(C)
                    if use_ns:
                        w_c = ops.whiten_ns(gram[0], 1.0 / (n - 1.0), add_identity=numpy_variant)
                    else:
                        c_e, c_v = ops.eigh_jacobi(gram, [1.0 / (n - 1.0)], add_identity=numpy_variant,   # util_wct.py:143: +I on content only
                                                   early_stop=self._early())
                        c_e, c_v = c_e[0], c_v[0]
                    mark(s, "eig")
                    (s_mean, s_e, s_v), ev = style_res[s]
                    if ev is not None:
                        main.wait_event(ev)
                    if use_ns:
                        m, b, mc = ops.wct_matrix_w(w_c, c_mean, s_e, s_v, s_mean, self.tau, float(alpha))
                    else:
                        m, b, mc = ops.wct_matrix(c_e, c_v, c_mean, s_e, s_v, s_mean, self.tau, float(alpha), self._keep(C), self._keep(C))
                    if self.fold_into_decoder:
                        L0 = getattr(dec, dec.layers[0]["name"])
                        w, bb = ops.fold_wct_into_conv(L0.weight.detach().contiguous(), L0.bias.detach().contiguous(), m, b, mc)
                        img = dec.forward_p4(c4 if c8 is None else c8, first_override=(w, bb))
                        del c8
                    else:
                        cs4 = ops.wct_apply(c4, m, b, mc, round_tf32=dec.first_layer_needs_tf32_input())
                        del c4
                        img = dec.forward_p4(cs4)
                    mark(s, "dec")
            if not torch.cuda.is_current_stream_capturing():
                img.record_stream(cur)
        cur.wait_stream(main)
        cur.wait_stream(side)
        return img

    def _weights_fingerprint(self, stages):
        """changes whenever a weight tensor of the nets on the path is replaced or updated in place (load_state_dict,
        load_npz_into, optimizer steps): a captured graph replays the PACKED weights of capture time, so it must not outlive them"""
        fp = []
        for s in stages:
            for net in (getattr(self, "e%d" % s), getattr(self, "d%d" % s)):
                fp.append(tuple((p.data_ptr(), p._version) for p in net.parameters()))
        return hash(tuple(fp))

    @torch.no_grad()
    def _stylize_graph(self, content, style, alpha, num_run, stages, style_cache=None):
        """Replay (capture on first use) a CUDA graph of the two-stream schedule for this input shape: ~500 kernel launches
        become one graph launch, which removes the CPU launch cost that dominates small / medium images.
        Device inputs are copied into static buffers; PINNED host inputs are captured in place (their H2D copies become
        graph nodes on the two branch streams, so the style branch starts while the content image is still in flight) --
        the graph is then keyed on the host buffer addresses."""
        if style_cache is not None:
            style = content[..., :1, :1]       # unused placeholder with a stable shape
        host = (not content.is_cuda) and (not style.is_cuda) and content.is_pinned() and style.is_pinned()
        key = (tuple(content.shape), tuple(style.shape), float(alpha), int(num_run), tuple(stages), nets.get_precision(),
               bool(self.fold_into_decoder), bool(getattr(self.args, "numpy", False)), self.num_eig, self.rat_eig, float(self.tau), self._early(), self.whiten_solver,
               (content.data_ptr(), style.data_ptr()) if host else None, id(style_cache) if style_cache is not None else None,
               torch.cuda.current_device(), bool(self.fast_stats), bool(self.fast_stats_h2), self._weights_fingerprint(stages))
        ent = self._graphs.get(key)
        if ent is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            if host:
                sc, ss = content, style
            else:
                sc = torch.empty(content.shape, dtype=torch.float32, device=dev)
                ss = torch.empty(style.shape, dtype=torch.float32, device=dev)
                sc.copy_(content, non_blocking=True)
                ss.copy_(style, non_blocking=True)
            self._stylize_two_streams(sc, ss, alpha, num_run, stages, style_cache)   # eager warm-up: packs weights, sets attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                n0 = ops.launches()
                with torch.cuda.graph(graph):
                    out = self._stylize_two_streams(sc, ss, alpha, num_run, stages, style_cache)
                ent = (graph, sc, ss, out, ops.launches() - n0, style_cache)
            except Exception as e:            # capture not possible on this setup: stay eager for this shape
                print("wct-b200: CUDA graph capture failed (%s); running eagerly" % (e,))
                torch.cuda.synchronize()
                ent = (None, None, None, None, 0, None)
            self._graphs[key] = ent
            while len(self._graphs) > max(1, int(self.max_graphs)):     # a folder of differently sized images must not
                self._graphs.popitem(last=False)                        # accumulate one private memory pool per shape
        else:
            self._graphs.move_to_end(key)
        graph, sc, ss, out, nlaunch, _keepalive = ent
        if graph is None:
            return self._stylize_two_streams(content, style, alpha, num_run, stages, style_cache)
        if not host:
            sc.copy_(content, non_blocking=True)
            ss.copy_(style, non_blocking=True)
        graph.replay()
        ops.add_launches(nlaunch)
        return out.clone()

    def pipeline(self, alpha=1.0, num_run=1, stages=(5, 4, 3, 2, 1), depth=2):
        """Throughput mode for a sequence of pinned host pairs: `pipeline.StylizePipeline` around stylize() -- the upload of pair
        i+1 and the download of result i-1 overlap the five stages of pair i (WCT.py:109-131 runs them back to back)."""
        from .pipeline import StylizePipeline
        return StylizePipeline(lambda c, s: self.stylize(c, s, alpha=alpha, num_run=num_run, stages=stages), depth=depth)

    @torch.no_grad()
    def stylize(self, content, style, alpha=1.0, num_run=1, stages=(5, 4, 3, 2, 1), style_cache=None):
        """content, style: [1,3,H,W] fp32 (CUDA, or CPU -> copied up).  Returns the stylized image on the GPU,
        un-clamped like the reference (WCT.py:120-125)."""
        if self.dist is None and self.overlap_style:
            # host (pinned) inputs are copied up on the two branch streams, device inputs are used in place
            c = content.float()
            st = c[..., :1, :1] if style_cache is not None else style.float()   # with a cache the style image is not needed
            if self.use_graph and getattr(self, "timeline", None) is None:
                return self._stylize_graph(c, st, alpha, num_run, tuple(stages), style_cache)
            return self._stylize_two_streams(c, st, alpha, num_run, tuple(stages), style_cache)
        if style is None:
            raise WctbError("stylize(style_cache=...) needs the single-GPU overlapped path (dist is None and overlap_style=True); "
                            "pass the style image here")
        img = content.to("cuda", torch.float32)
        style = style.to("cuda", torch.float32)
        for _ in range(num_run):
            for s in stages:
                img = self.style_transfer_stage(s, img, style, alpha)
        return img
