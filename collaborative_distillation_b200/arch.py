"""Layer plans of the five encoder/decoder pairs (VGG-19 prefixes), table driven.

reference: model/model_cd.py:688-702 (16x encoder 5), :246-258 (16x decoder 5), :324 (encoder 1 is 3->24);
model/model_original.py:434-446 / :539-551 (unpruned).  Encoder k is the VGG-19 prefix ending at conv{k}_1;
decoder k mirrors it (conv{j}_1 maps width j -> width j-1, conv1_1 -> 3 channels) with nearest x2
upsampling where the encoder pooled, and a ReLU after every conv including the last (model_cd.py:293).
"""
from __future__ import annotations

VGG = ("conv11", "conv12", "P", "conv21", "conv22", "P", "conv31", "conv32", "conv33", "conv34", "P",
       "conv41", "conv42", "conv43", "conv44", "P", "conv51")
WIDTH = {"original": (64, 128, 256, 512, 512), "16x": (16, 32, 64, 128, 128)}
WIDTH["16x_kd2sd"] = WIDTH["16x"]
MODES = tuple(WIDTH)


def _width(mode, stage, level):
    if mode != "original" and stage == 1:
        return 24                       # SmallEncoder1_16x_aux / SmallDecoder1_16x (model_cd.py:324, :67)
    return WIDTH[mode][level - 1]


def encoder_layers(mode: str, stage: int):
    """-> list of dicts {name, cin, cout, pool_after}"""
    out, cin = [], 3
    for item in VGG:
        if item == "P":
            out[-1]["pool_after"] = True
            continue
        cout = _width(mode, stage, int(item[4]))
        out.append({"name": item, "cin": cin, "cout": cout, "pool_after": False})
        cin = cout
        if item == "conv%d1" % stage:
            break
    return out


def decoder_layers(mode: str, stage: int):
    """-> list of dicts {name, cin, cout, up_after} in execution order."""
    enc = encoder_layers(mode, stage)
    out = []
    for k in range(len(enc) - 1, -1, -1):
        e = enc[k]
        # the encoder pooled after layer k-1  <=>  the decoder upsamples after its layer k
        up = k > 0 and enc[k - 1]["pool_after"]
        out.append({"name": e["name"], "cin": e["cout"], "cout": e["cin"], "up_after": up})
    return out


def t7_indices(kind: str, stage: int):
    """child index (0-based) of every conv inside the Torch7 nn.Sequential of the WCT authors' files, derived from the
    module order [conv0] (pad conv relu)+ [pool|unpool]; equals the literal tables of model_original.py
    (e.g. Encoder5 :471-484 conv0:0 conv11:2 ... conv51:42, Decoder5 :561-573 conv51:1 ... conv11:41)."""
    out = {}
    if kind == "enc":
        out["conv0"], idx = 0, 1
        for L in encoder_layers("original", stage):
            out[L["name"]] = idx + 1
            idx += 4 if L["pool_after"] else 3
    else:
        idx = 0
        for L in decoder_layers("original", stage):
            out[L["name"]] = idx + 1
            idx += 4 if L["up_after"] else 3
    return out


def feature_channels(mode: str, stage: int) -> int:
    return encoder_layers(mode, stage)[-1]["cout"]


def feature_hw(stage: int, H: int, W: int):
    """floor-mode pooling: (H >> (stage-1), W >> (stage-1))  (SURVEY 8(a) note 3)"""
    return H >> (stage - 1), W >> (stage - 1)


# aux heads that exist in the reference classes (and the shipped .pth) but are never used by forward()
ENCODER_AUX = {"16x": {1: (("conv11_aux", 24, 64),),
                       2: (("conv11_aux", 16, 64), ("conv21_aux", 32, 128)),
                       3: (("conv11_aux", 16, 64), ("conv21_aux", 32, 128), ("conv31_aux", 64, 256)),
                       4: (("conv11_aux", 16, 64), ("conv21_aux", 32, 128), ("conv31_aux", 64, 256), ("conv41_aux", 128, 512)),
                       5: (("conv11_aux", 16, 64), ("conv21_aux", 32, 128), ("conv31_aux", 64, 256), ("conv41_aux", 128, 512),
                           ("conv51_aux", 128, 512))}}
DECODER_AUX_KD2SD = {1: (), 2: (("aux21", 16, 64),), 3: (("aux31", 32, 128), ("aux21", 16, 64)),
                     4: (("aux41", 64, 256), ("aux31", 32, 128), ("aux21", 16, 64)),
                     5: (("aux51", 128, 512), ("aux41", 64, 256), ("aux31", 32, 128), ("aux21", 16, 64))}
