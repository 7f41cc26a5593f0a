"""Drop-in encoder / decoder modules with the reference's class names and call signatures.

reference: model/model_cd.py (SmallEncoder{1..5}_16x_aux, SmallDecoder{1..5}_16x),
model/model_kd2sd.py (SmallDecoder{1..5}_16x_aux), model/model_original.py (Encoder{1..5}, Decoder{1..5}).
`Cls(model=None, fixed=False)`; `forward(x[1,3,H,W]) -> [1,C,h,w]` / `forward(y[1,C,h,w]) -> [1,3,H',W']`.
state_dict keys are the reference's (`conv0`, `conv11`, ..., `conv*_aux`, `aux*`), so the shipped .pth load unchanged.

Unlike the reference, forward() runs hand-written sm_100a kernels (libwctb.so) and therefore needs CUDA
tensors; a CPU tensor raises (no silent fallback).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import arch, ops
from ._lib import WctbError

LAST_TC = True        # TF32 mode: an unfused last decoder layer (C -> 3) runs on the tensor cores with N padded to 16
FUSE_TAIL = True      # [x2 upsample +] conv12 + conv11 of the decoders in one kernel when the TF32 engine is active
HEAD_TC = True        # 16x nets: conv11 of the fused head on the tensor cores too (else FFMA producers)
FUSE_HEAD = True      # conv11+conv12(+pool) in one kernel when the TF32 engine is active
# "h2"  (default): tcgen05 kind::f16 on fp16 hi/lo operand pairs -- fp32-accurate tensor-core convs (csrc/conv_h2.cu)
# "tf32": single-pass tcgen05 TF32 engine (csrc/conv_umma.cu): 10-bit operands, does NOT meet the 3e-3 RMS contract on
#         noise-like inputs; kept as the fast / lossy option and as an A/B
# "fp32": CUDA-core fp32 everywhere (the GPU-side reference of the other two)
_PRECISION = os.environ.get("WCTB_PRECISION", "h2")
PRECISIONS = ("h2", "tf32", "fp32")


def set_precision(p: str):
    global _PRECISION
    if p not in PRECISIONS:
        raise ValueError("precision must be one of %s" % (PRECISIONS,))
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def _load_state(module: nn.Module, model):
    """model_cd.py:72-76 / model_original.py:24-30: accept {'model': sd} or a bare state_dict."""
    if not model:
        return
    ext = os.path.splitext(model)[1]
    if ext == ".t7":                       # model_original.py:25-29: load_lua + load_param per sequential child
        from . import t7
        seq = t7.load_t7(model)
        for name, idx in arch.t7_indices(module.KIND, module.STAGE).items():
            t7.load_param_from_t7(seq, idx, getattr(module, name))
        print("load model '%s' successfully" % model)
        return
    assert ext == ".pth", "weights must be .pth or .t7 (model_original.py:24)"
    sd = torch.load(model, map_location="cpu")
    if isinstance(sd, dict) and "model" in sd and not torch.is_tensor(sd["model"]):
        sd = sd["model"]
    module.load_state_dict(sd)
    print("load model '%s' successfully" % model)


class _Net(nn.Module):
    MODE = "16x"
    STAGE = 5
    KIND = "enc"
    AUX = ()

    def __init__(self, model=None, fixed=False):
        super().__init__()
        self.fixed = fixed
        if self.KIND == "enc":
            self.layers = arch.encoder_layers(self.MODE, self.STAGE)
            self.conv0 = nn.Conv2d(3, 3, 1, 1, 0)
            if self.MODE == "original" and self.STAGE == 5:   # model_original.py:427-433
                with torch.no_grad():
                    self.conv0.weight.copy_(torch.tensor([[0., 0, 255], [0, 255., 0], [255., 0, 0]]).view(3, 3, 1, 1))
                    self.conv0.bias.copy_(torch.tensor([-103.939, -116.779, -123.68]))
        else:
            self.layers = arch.decoder_layers(self.MODE, self.STAGE)
        for L in self.layers:
            setattr(self, L["name"], nn.Conv2d(L["cin"], L["cout"], 3, 1, 0))
        for name, cin, cout in self.AUX:
            setattr(self, name, nn.Conv2d(cin, cout, 1, 1, 0))
        _load_state(self, model)
        if fixed:
            for p in self.parameters():
                p.requires_grad = False
        self._pack_cache = {}

    # ---------------------------------------------------------------- packing
    def _cache_key(self, precision):
        ps = [getattr(self, L["name"]).weight for L in self.layers]
        return (precision, ps[0].device, tuple((p._version, p.data_ptr()) for p in ps))

    def _engine(self, L, precision):
        if precision == "tf32" and ops.tf32_supported(L["cin"], L["cout"]):
            return ops.ENGINE_TF32
        return ops.ENGINE_FP32

    def _pack_layer(self, idx, precision, w=None, b=None):
        L = self.layers[idx]
        conv = getattr(self, L["name"])
        w = conv.weight.detach() if w is None else w
        b = conv.bias.detach() if b is None else b
        first = self.KIND == "enc" and idx == 0
        last = self.KIND == "dec" and idx == len(self.layers) - 1
        if first:   # fold conv0 (1x1) into conv11: exact under reflection padding (SURVEY 8(a) note 1)
            w0 = self.conv0.weight.detach().double().view(3, 3)
            b0 = self.conv0.bias.detach().double()
            wd = w.double()
            b = (b.double() + torch.einsum("ojyx,j->o", wd, b0)).float()
            w = torch.einsum("ojyx,ji->oiyx", wd, w0).float()
        if precision == "h2":
            return self._pack_layer_h2(idx, L, w, b, first, last)
        engine = ops.ENGINE_FP32 if (first or last) else self._engine(L, precision)
        out = {"w": ops.pack_weights(w.contiguous(), engine), "b": b.contiguous().float(), "engine": engine}
        if last and precision == "tf32" and LAST_TC and ops.tf32_supported(L["cin"], 16):
            wp = torch.zeros(16, w.shape[1], 3, 3, device=w.device, dtype=torch.float32)
            wp[:3] = w
            bp = torch.zeros(16, device=w.device, dtype=torch.float32)
            bp[:3] = b.float()
            out["w_last_tc"], out["b_last_tc"] = ops.pack_weights(wp, ops.ENGINE_TF32), bp
        if first and precision == "tf32" and L["cout"] == 16:
            out["w_tc"] = ops.pack_head_tc_weights(w)       # conv11 on the tensor cores (fused head of the 16x nets)
        return out

    def _pack_layer_h2(self, idx, L, w, b, first, last):
        """h2 engine: first layer (3 -> C) stays fp32 FFMA; a C -> 3 last layer is zero-padded to 16 outputs.  The layers
        that the fused head / tail kernels cover (static weights) also get their dx-stacked tensor-core tiles."""
        if first:
            out = {"w": ops.pack_weights(w.contiguous(), ops.ENGINE_FP32), "b": b.contiguous().float(), "engine": ops.ENGINE_H2}
            if L["cout"] == 16:
                out["w11_h2"], out["inv_s11"] = ops.pack_head_h2_w11(w.contiguous())
            return out
        cin, cout = w.shape[1], w.shape[0]
        if last:
            wp = torch.zeros(16, cin, 3, 3, device=w.device, dtype=torch.float32)
            wp[:cout] = w
            bp = torch.zeros(16, device=w.device, dtype=torch.float32)
            bp[:cout] = b.float()
            w, b, cout = wp, bp, 16
        if not ops.h2_supported(cin, cout):
            raise WctbError("h2 engine: unsupported layer %d -> %d" % (cin, cout))
        wh, ws = ops.pack_weights_h2(w.contiguous())
        out = {"w": wh, "ws": ws, "b": b.contiguous().float(), "engine": ops.ENGINE_H2, "cin": cin, "cout": cout}
        n = len(self.layers)
        in_head = self.KIND == "enc" and idx == 1 and cin == 16 and cout == 16 and L["pool_after"]
        in_tail = self.KIND == "dec" and n >= 3 and idx >= n - 2 and cin == 16
        if in_head or in_tail:
            out["w_dx"], out["inv_s_dx"] = ops.pack_dx_h2(w.contiguous())
        return out

    def packed(self, precision=None):
        precision = precision or _PRECISION
        key = self._cache_key(precision)
        if self._pack_cache.get("key") != key:
            layers = [self._pack_layer(i, precision) for i in range(len(self.layers))]
            n = len(layers)
            if (self.KIND == "dec" and n >= 3 and layers[n - 2]["engine"] == ops.ENGINE_TF32
                    and ops.conv_tail_supported(self.layers[n - 2]["cin"], self.layers[n - 2]["cout"])):
                # fused tail runs conv11 on the tensor cores too: zero-pad its 3 output channels to 16, pack as TF32
                w = getattr(self, self.layers[n - 1]["name"]).weight.detach()
                wp = torch.zeros(16, w.shape[1], 3, 3, device=w.device, dtype=torch.float32)
                wp[:3] = w
                layers[n - 1]["w_tail"] = ops.pack_weights(wp, ops.ENGINE_TF32)
            self._pack_cache = {"key": key, "layers": layers}
        return self._pack_cache["layers"]

    @staticmethod
    def _check_cuda(x):
        if not x.is_cuda:
            raise WctbError("this module runs sm_100a CUDA kernels: input must be a CUDA tensor (no CPU fallback)")


class _Encoder(_Net):
    KIND = "enc"

    def forward_p4(self, x, precision=None, round_output=False):
        """x [1,3,H,W] (or [3,H,W]) CUDA fp32 -> P4 feature [C/4,h,w,4].  round_output: store the feature TF32-rounded
        (rna) because a tensor-core layer consumes it directly (WCT matrix folded into the decoder's first conv)."""
        precision = precision or _PRECISION
        if precision == "h2":
            return self.forward_feat(x, want_h8=False)[0]
        self._check_cuda(x)
        pk = self.packed(precision)
        x = x.detach().contiguous().float()
        H, W = x.shape[-2:]
        n = len(self.layers)
        sh, sw = arch.feature_hw(self.STAGE, H, W)
        if sh < 2 or sw < 2:
            raise WctbError("input %dx%d too small for stage %d (ReflectionPad2d needs >=2 px at the deepest level)" % (H, W, self.STAGE))
        nxt = lambda i: pk[i + 1]["engine"] == ops.ENGINE_TF32 if i + 1 < n else bool(round_output)
        L0 = self.layers[0]
        if (FUSE_HEAD and n >= 2 and pk[1]["engine"] == ops.ENGINE_TF32
                and ops.conv_head_supported(L0["cout"], self.layers[1]["cout"])):
            L1 = self.layers[1]
            epi = ops.EPI_POOL2 if L1["pool_after"] else ops.EPI_NONE
            if HEAD_TC and "w_tc" in pk[0] and L1["cout"] == 16:
                y = ops.conv_head_tc(x, pk[0]["w_tc"], pk[0]["b"], pk[1]["w"], pk[1]["b"], epi, nxt(1))
            else:
                y = ops.conv_head(x, pk[0]["w"], pk[0]["b"], pk[1]["w"], pk[1]["b"], L0["cout"], L1["cout"], epi, nxt(1))
            first = 2
        else:
            y = ops.conv3x3_first(x, pk[0]["w"], pk[0]["b"], L0["cout"], nxt(0))
            first = 1
        for i in range(first, n):
            L = self.layers[i]
            epi = ops.EPI_POOL2 if L["pool_after"] else ops.EPI_NONE
            y = ops.conv3x3_p4(y, pk[i]["w"], pk[i]["b"], L["cout"], epi, nxt(i), pk[i]["engine"])
        return y

    def forward_feat(self, x, want_h8=True, want_p4=True, precision=None):
        """x [1,3,H,W] CUDA fp32 -> (P4 fp32 feature | None, H8 feature | None).  The h2 engine's last layer can write both
        forms at once: fp32 P4 for the statistics kernels, H8 for the decoder's (WCT-folded) first conv.  With the other
        engines the P4 tensor serves both purposes and the second element is None."""
        precision = precision or _PRECISION
        if precision != "h2":
            return self.forward_p4(x, precision, round_output=want_h8), None
        self._check_cuda(x)
        pk = self.packed(precision)
        x = x.detach().contiguous().float()
        H, W = x.shape[-2:]
        n = len(self.layers)
        sh, sw = arch.feature_hw(self.STAGE, H, W)
        if sh < 2 or sw < 2:
            raise WctbError("input %dx%d too small for stage %d (ReflectionPad2d needs >=2 px at the deepest level)" % (H, W, self.STAGE))
        L0 = self.layers[0]
        last = n == 1
        first = 1
        if FUSE_HEAD and n >= 3 and "w11_h2" in pk[0] and "w_dx" in pk[1]:
            # conv11 + conv12 + pool in one kernel (16x nets, stages 2..5)
            y8, y4 = ops.conv_head_h2(x, pk[0]["w11_h2"], pk[0]["inv_s11"], pk[0]["b"], pk[1]["w_dx"], pk[1]["inv_s_dx"], pk[1]["b"]), None
            first = 2
        else:
            y8, y4 = ops.conv3x3_first_h2(x, pk[0]["w"], pk[0]["b"], L0["cout"], out_h8=(not last) or want_h8, out_p4=last and want_p4)
        for i in range(first, n):
            L = self.layers[i]
            last = i == n - 1
            epi = ops.EPI_POOL2 if L["pool_after"] else ops.EPI_NONE
            y8, y4 = ops.conv3x3_h2(y8, pk[i]["w"], pk[i]["ws"], pk[i]["b"], L["cin"], L["cout"], epi,
                                    out_h8=(not last) or want_h8, out_p4=last and want_p4)
        return y4, y8

    def forward(self, x):
        return ops.p4_to_nchw(self.forward_p4(x))


class _Decoder(_Net):
    KIND = "dec"

    def forward_p4(self, y, precision=None, first_override=None, tail_shard=None):
        """y P4 [C/4,h,w,4] -> image [1,3,H,W].  first_override=(w_oihw, bias) replaces the first conv's
        parameters (used to fold the WCT matrix into it).  tail_shard: see ops.conv_tail_h2 (strip-sharded output written
        by the fused tail kernel; ignored -- a plain image is returned -- when this decoder has no fused h2 tail)."""
        self._check_cuda(y)
        precision = precision or _PRECISION
        pk = list(self.packed(precision))
        if first_override is not None:
            pk[0] = self._pack_layer(0, precision, first_override[0], first_override[1])
        n = len(self.layers)
        if precision == "h2":
            return self._forward_h2(y, pk, tail_shard)
        if y.dtype == torch.float16:
            raise WctbError("an H8 feature needs the h2 engine")
        if y.shape[1] < 2 or y.shape[2] < 2:
            raise WctbError("feature map too small for ReflectionPad2d(1)")
        nxt = lambda i: (pk[i + 1]["engine"] == ops.ENGINE_TF32 or "w_last_tc" in pk[i + 1]) if i + 1 < n else False
        # fused tail: [x2] conv12 + conv11 in one kernel (TF32 engine, 16-channel nets)
        fuse_tail = FUSE_TAIL and n >= 3 and "w_tail" in pk[n - 1]
        last_plain = n - 2 if fuse_tail else n - 1
        up_in = False
        for i in range(last_plain):
            L = self.layers[i]
            epi = ops.EPI_UP2 if L["up_after"] else ops.EPI_NONE
            if fuse_tail and i == n - 3 and L["up_after"]:
                epi, up_in = ops.EPI_NONE, True          # the tail kernel upsamples while it loads
            y = ops.conv3x3_p4(y, pk[i]["w"], pk[i]["b"], L["cout"], epi, nxt(i), pk[i]["engine"])
        if fuse_tail:
            return ops.conv_tail(y, pk[n - 2]["w"], pk[n - 2]["b"], pk[n - 1]["w_tail"], pk[n - 1]["b"], up_in)
        if "w_last_tc" in pk[n - 1]:      # TF32 mode: the C -> 3 layer on the tensor cores (3 outputs zero-padded to 16)
            return ops.conv3x3_p4(y, pk[n - 1]["w_last_tc"], pk[n - 1]["b_last_tc"], 16, ops.EPI_NCHW3, False, ops.ENGINE_TF32)
        return ops.conv3x3_last(y, pk[n - 1]["w"], pk[n - 1]["b"])

    def _forward_h2(self, y, pk, tail_shard=None):
        """h2 engine: y is an H8 feature (or fp32 P4, converted) -> image [1,3,H,W]"""
        if y.dtype != torch.float16:
            y = ops.p4_to_h8(y)
        if y.shape[2] < 2 or y.shape[3] < 2:
            raise WctbError("feature map too small for ReflectionPad2d(1)")
        n = len(self.layers)
        fuse_tail = FUSE_TAIL and n >= 3 and "w_dx" in pk[n - 2] and "w_dx" in pk[n - 1]
        for i in range(n - 2 if fuse_tail else n - 1):
            L = self.layers[i]
            epi = ops.EPI_UP2 if L["up_after"] else ops.EPI_NONE
            if fuse_tail and i == n - 3:
                epi = ops.EPI_NONE                       # the tail kernel upsamples while it loads
            y, _ = ops.conv3x3_h2(y, pk[i]["w"], pk[i]["ws"], pk[i]["b"], L["cin"], L["cout"], epi)
        if fuse_tail:
            return ops.conv_tail_h2(y, pk[n - 2]["w_dx"], pk[n - 2]["inv_s_dx"], pk[n - 2]["b"], pk[n - 1]["w_dx"], pk[n - 1]["inv_s_dx"],
                                    pk[n - 1]["b"], bool(self.layers[n - 3]["up_after"]), shard=tail_shard)
        L = self.layers[n - 1]
        return ops.conv3x3_h2(y, pk[n - 1]["w"], pk[n - 1]["ws"], pk[n - 1]["b"], L["cin"], 16, ops.EPI_NCHW3)[1]

    def first_layer_needs_tf32_input(self, precision=None):
        if (precision or _PRECISION) == "h2":
            return False
        pk = self.packed(precision or _PRECISION)
        if len(pk) == 1:
            return "w_last_tc" in pk[0]
        return pk[0]["engine"] == ops.ENGINE_TF32

    def forward(self, y):
        self._check_cuda(y)
        return self.forward_p4(ops.nchw_to_p4(y.detach().contiguous().float(), self.first_layer_needs_tf32_input()))


def _make(name, base, mode, stage, aux=()):
    return type(name, (base,), {"MODE": mode, "STAGE": stage, "AUX": tuple(aux), "__doc__":
                                "%s stage %d, mode %s (see module docstring)" % (base.KIND, stage, mode)})


_g = globals()
for _k in range(1, 6):
    _g["Encoder%d" % _k] = _make("Encoder%d" % _k, _Encoder, "original", _k)
    _g["Decoder%d" % _k] = _make("Decoder%d" % _k, _Decoder, "original", _k)
    _g["SmallEncoder%d_16x_aux" % _k] = _make("SmallEncoder%d_16x_aux" % _k, _Encoder, "16x", _k, arch.ENCODER_AUX["16x"][_k])
    _g["SmallDecoder%d_16x" % _k] = _make("SmallDecoder%d_16x" % _k, _Decoder, "16x", _k)
    _g["SmallDecoder%d_16x_aux" % _k] = _make("SmallDecoder%d_16x_aux" % _k, _Decoder, "16x_kd2sd", _k, arch.DECODER_AUX_KD2SD[_k])

ENCODERS = {"original": [_g["Encoder%d" % k] for k in range(1, 6)],
            "16x": [_g["SmallEncoder%d_16x_aux" % k] for k in range(1, 6)]}
ENCODERS["16x_kd2sd"] = ENCODERS["16x"]
DECODERS = {"original": [_g["Decoder%d" % k] for k in range(1, 6)],
            "16x": [_g["SmallDecoder%d_16x" % k] for k in range(1, 6)],
            "16x_kd2sd": [_g["SmallDecoder%d_16x_aux" % k] for k in range(1, 6)]}
