"""CPU-only checks of the host side: plans, class surface, state_dict compatibility, C-ABI exports."""
import ctypes
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import collaborative_distillation_b200 as P
from collaborative_distillation_b200 import _lib, arch, nets
from oracle import wct_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["16x", "original", "16x_kd2sd"])
def test_plans_agree_with_oracle(mode):
    for s in range(1, 6):
        enc = [(L["name"], L["cin"], L["cout"]) for L in arch.encoder_layers(mode, s)]
        assert enc == [p for p in O.encoder_plan(mode, s) if p != "P"]
        dec = [(L["name"], L["cin"], L["cout"]) for L in arch.decoder_layers(mode, s)]
        assert dec == [p for p in O.decoder_plan(mode, s) if p != "U"]
        # pool / upsample positions
        oplan = O.encoder_plan(mode, s)
        pools = [oplan[i - 1][0] for i, p in enumerate(oplan) if p == "P"]
        assert [L["name"] for L in arch.encoder_layers(mode, s) if L["pool_after"]] == pools
        dplan = O.decoder_plan(mode, s)
        ups = [dplan[i - 1][0] for i, p in enumerate(dplan) if p == "U"]
        assert [L["name"] for L in arch.decoder_layers(mode, s) if L["up_after"]] == ups


def test_floor_pool_bookkeeping():
    # SURVEY 8(a) note 3: 420 -> 416 at stage 5; 591x800 -> 576x800
    for (H, W) in [(420, 420), (591, 800), (84, 100)]:
        h, w = arch.feature_hw(5, H, W)
        x = torch.zeros(1, 1, H, W)
        for _ in range(4):
            x = torch.nn.functional.max_pool2d(x, 2, 2)
        assert (h, w) == tuple(x.shape[-2:])
    assert arch.feature_hw(5, 420, 420) == (26, 26) and 26 * 16 == 416


def test_reference_class_names_and_signatures():
    for k in range(1, 6):
        for n in ("Encoder%d", "Decoder%d", "SmallEncoder%d_16x_aux", "SmallDecoder%d_16x", "SmallDecoder%d_16x_aux"):
            cls = getattr(nets, n % k)
            m = cls(None, False)
            assert isinstance(m, torch.nn.Module)
    e5 = nets.Encoder5()
    assert torch.equal(e5.conv0.bias.detach(), torch.tensor([-103.939, -116.779, -123.68]))  # model_original.py:431-433


def test_shipped_state_dict_keys_load_strict(golden_dir):
    # the golden npz holds every non-aux tensor of the shipped .pth; aux heads exist as modules so strict loads work
    z = np.load(os.path.join(golden_dir, "weights_16x.npz"))
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    have = set()
    for s in range(1, 6):
        for tag in ("e", "d"):
            for k in getattr(w, "%s%d" % (tag, s)).state_dict():
                have.add("%s%d.%s" % (tag, s, k))
    assert set(z.files) <= have
    assert any("conv51_aux" in k for k in have)
    P.weights.load_npz_into(w, os.path.join(golden_dir, "weights_16x.npz")) if hasattr(P, "weights") else None


def test_wrong_mode_and_cpu_input_fail_loudly():
    with pytest.raises(ValueError):
        P.WCT(SimpleNamespace(mode="bogus", numpy=False))
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    with pytest.raises(P.WctbError):
        w.e1(torch.rand(1, 3, 16, 16))          # CPU tensor: no fallback
    with pytest.raises(P.WctbError):
        w.d1(torch.rand(1, 24, 16, 16))


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wctb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wctb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libwctb.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES) | {"wctb_error_string"}
    assert _lib.load().wctb_abi_version() == 1
    assert _lib.load().wctb_error_string(-2) == b"unsupported configuration"


def test_cli_flag_surface_matches_reference():
    """WCT.py:15-34 flag names/defaults + mode->weight path tables (:36-70)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("wct_cli", os.path.join(ROOT, "PytorchWCT", "WCT.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    a = cli.parse(["--mode", "16x", "--UHD", "--alpha", "0.6", "--num_run", "2", "--debug"])
    assert a.mode == "16x" and a.UHD and a.alpha == 0.6 and a.num_run == 2 and a.debug and not a.numpy and not a.synthesis
    assert a.e5 == "../trained_models/wct_se_16x_new/5SE.pth" and a.d1 == "../trained_models/wct_se_16x_new_sd/1SD.pth"
    assert (a.contentPath, a.stylePath, a.UHD_contentPath, a.UHD_stylePath, a.outf) == (
        "content", "style", "content/UHD_content", "style/UHD_style", "stylized_results")
    assert a.content_size == 0 and a.style_size == 0 and a.picked_content_mark == "." and a.texturePath == "style/texture"
    d = cli.parse([])
    assert d.mode is None and d.e3.endswith("original_wct_models/vgg_normalised_conv3_1.t7") and d.alpha == 1
    k = cli.parse(["--mode", "16x_kd2sd"])
    assert k.d4 == "../trained_models/wct_se_16x_new_sd_kd2sd/4SD.pth"


def test_reference_import_paths():
    import importlib
    sys_path = os.path.join(ROOT, "PytorchWCT")
    import sys
    sys.path.insert(0, sys_path)
    try:
        m = importlib.import_module("model.model_cd")
        assert hasattr(m, "SmallEncoder5_16x_aux") and hasattr(m, "SmallDecoder1_16x")
        assert hasattr(importlib.import_module("model.model_original"), "Decoder4")
        assert hasattr(importlib.import_module("model.model_kd2sd"), "SmallDecoder3_16x_aux")
        assert importlib.import_module("util_wct").WCT is P.WCT
    finally:
        sys.path.remove(sys_path)
