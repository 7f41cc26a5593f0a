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
    assert declared == set(_lib.SIGNATURES) | {"wctb_error_string", "wctb_workspace_doubles", "wctb_h2_packed_halves"}
    lib2 = _lib.load()
    assert lib2.wctb_workspace_doubles(_lib.WS_EIGH, 128, 2) == 2 * 128 * 128 + 16
    assert lib2.wctb_workspace_doubles(_lib.WS_WCT_MATRIX, 64, 1) == 3 * 64 * 64 + 8
    assert lib2.wctb_workspace_doubles(_lib.WS_WHITEN_NS, 128, 1) == 8 * 128 * 128 + 8
    assert lib2.wctb_workspace_doubles(7, 64, 1) == -1 and lib2.wctb_workspace_doubles(_lib.WS_EIGH, 0, 1) == -1
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


# ---------------------------------------------------------------------------------------------------------------------
# Torch7 weights of --mode original (model_original.py:25-29, utils.py:64-67)
_T7_TABLE = {  # literal indices of the reference's load_param calls (model_original.py)
    ("enc", 1): {"conv0": 0, "conv11": 2},
    ("dec", 1): {"conv11": 1},
    ("enc", 2): {"conv0": 0, "conv11": 2, "conv12": 5, "conv21": 9},
    ("dec", 2): {"conv21": 1, "conv12": 5, "conv11": 8},
    ("enc", 3): {"conv0": 0, "conv11": 2, "conv12": 5, "conv21": 9, "conv22": 12, "conv31": 16},
    ("dec", 3): {"conv31": 1, "conv22": 5, "conv21": 8, "conv12": 12, "conv11": 15},
    ("enc", 4): {"conv0": 0, "conv11": 2, "conv12": 5, "conv21": 9, "conv22": 12, "conv31": 16, "conv32": 19, "conv33": 22,
                 "conv34": 25, "conv41": 29},
    ("dec", 4): {"conv41": 1, "conv34": 5, "conv33": 8, "conv32": 11, "conv31": 14, "conv22": 18, "conv21": 21, "conv12": 25,
                 "conv11": 28},
    ("enc", 5): {"conv0": 0, "conv11": 2, "conv12": 5, "conv21": 9, "conv22": 12, "conv31": 16, "conv32": 19, "conv33": 22,
                 "conv34": 25, "conv41": 29, "conv42": 32, "conv43": 35, "conv44": 38, "conv51": 42},
    ("dec", 5): {"conv51": 1, "conv44": 5, "conv43": 8, "conv42": 11, "conv41": 14, "conv34": 18, "conv33": 21, "conv32": 24,
                 "conv31": 27, "conv22": 31, "conv21": 34, "conv12": 38, "conv11": 41},
}


def test_t7_indices_match_reference_tables():
    from collaborative_distillation_b200 import arch
    for (kind, stage), want in _T7_TABLE.items():
        assert arch.t7_indices(kind, stage) == want, (kind, stage)


def _fake_sequential(kind, stage, seed):
    """an nn.Sequential laid out like the WCT authors' files: [conv0] (pad conv relu)+ [pool|unpool]"""
    from collaborative_distillation_b200 import arch, t7
    g = torch.Generator().manual_seed(seed)
    mods, convs = [], {}

    def conv(name, cin, cout, k):
        w, b = torch.randn(cout, cin, k, k, generator=g), torch.randn(cout, generator=g)
        convs[name] = (w, b)
        return t7.T7Object("nn.SpatialConvolution", {"weight": w, "bias": b, "nInputPlane": cin, "nOutputPlane": cout,
                                                      "kW": k, "kH": k, "train": False, "gradWeight": None})
    if kind == "enc":
        mods.append(conv("conv0", 3, 3, 1))
        layers = arch.encoder_layers("original", stage)
    else:
        layers = arch.decoder_layers("original", stage)
    for L in layers:
        mods += [t7.T7Object("nn.SpatialReflectionPadding", {"pad_l": 1, "pad_r": 1, "pad_t": 1, "pad_b": 1}),
                 conv(L["name"], L["cin"], L["cout"], 3), t7.T7Object("nn.ReLU", {"inplace": True, "threshold": 0})]
        if L.get("pool_after"):
            mods.append(t7.T7Object("nn.SpatialMaxPooling", {"kW": 2, "kH": 2, "dW": 2, "dH": 2, "ceil_mode": False}))
        if L.get("up_after"):
            mods.append(t7.T7Object("nn.SpatialUpSamplingNearest", {"scale_factor": 2}))
    return t7.T7Object("nn.Sequential", {"modules": mods, "train": False}), convs


@pytest.mark.parametrize("kind,stage", [("enc", 1), ("enc", 3), ("dec", 3), ("dec", 2)])
def test_t7_weights_load_into_original_modules(tmp_path, kind, stage):
    from collaborative_distillation_b200 import nets, t7
    seq, convs = _fake_sequential(kind, stage, seed=stage)
    path = str(tmp_path / ("%s%d.t7" % (kind, stage)))
    t7.save_t7(path, seq)
    back = t7.load_t7(path)
    assert back.torch_typename == "nn.Sequential" and len(back.modules) == len(seq.modules)
    assert back.get(0).torch_typename == seq.get(0).torch_typename
    cls = getattr(nets, ("Encoder%d" if kind == "enc" else "Decoder%d") % stage)
    m = cls(path)
    for name, (w, b) in convs.items():
        assert torch.equal(getattr(m, name).weight.detach(), w), name
        assert torch.equal(getattr(m, name).bias.detach(), b), name


def test_t7_reader_shared_storage_and_views(tmp_path):
    """tensors that share one storage (weight views) and repeated references resolve through the ref-index memo"""
    import struct
    from collaborative_distillation_b200 import t7
    p = str(tmp_path / "v.t7")
    base = torch.arange(12, dtype=torch.float32)
    with open(p, "wb") as f:
        w = t7._Writer(f)
        # table {a = tensor[2,3] @offset 1, b = transposed view [3,2] of the SAME storage (by reference), c = a (by reference)}
        w.w("ii", 3, 1); w.w("i", 3)
        w.obj("a")
        w.w("ii", 4, 2); w.string("V 1"); w.string("torch.FloatTensor")
        w.w("i", 2); w.w("qq", 2, 3); w.w("qq", 3, 1); w.w("q", 1)
        w.w("ii", 4, 3); w.string("V 1"); w.string("torch.FloatStorage"); w.w("q", 12); f.write(base.numpy().tobytes())
        w.obj("b")
        w.w("ii", 4, 4); w.string("V 1"); w.string("torch.FloatTensor")
        w.w("i", 2); w.w("qq", 3, 2); w.w("qq", 1, 3); w.w("q", 7)
        w.w("ii", 4, 3)                                                   # storage by reference
        w.obj("c")
        w.w("ii", 4, 2)                                                   # tensor by reference
    out = t7.load_t7(p)
    assert torch.equal(out["a"], base[:6].view(2, 3))
    assert torch.equal(out["b"], base[6:12].view(2, 3).t())
    assert out["c"] is out["a"]
    with open(p, "rb") as f:
        blob = f.read()
    with open(p, "wb") as f:
        f.write(blob[:-5])
    with pytest.raises(EOFError):
        t7.load_t7(p)


# ---------------------------------------------------------------------------------------------------------------------
# the owner/visitor column schedule of the shared-memory eigensolver (csrc/wct_transform.cu: jacobi_chol_sweeps)
@pytest.mark.parametrize("k", [2, 4, 24, 28, 52, 60, 100, 128])
def test_jacobi_owner_visitor_schedule(k):
    """Replays the ring bookkeeping of the kernel on the host: every group keeps one column in registers ('owner'), its
    partner ('visitor') goes through shared memory.  Checks that every unordered pair meets exactly once per sweep, that no
    group ever reads a shared-memory copy that is stale (owned by someone's registers), and the hand-over rule."""
    m, h, ng = k - 1, (k - 2) // 2, k // 2
    valid = [True] * k                       # shared-memory copy of column c is current
    own = [k - 1 if g == 0 else g for g in range(ng)]
    for c in own:
        assert valid[c]
        valid[c] = False
    seen = set()
    for r in range(m):
        visitors, handing = [], []
        for g in range(ng):
            c = own[g]
            if g == 0:
                q = r
            else:
                q = 2 * r - c
                q = q + m if q < 0 else (q - m if q >= m else q)
            assert q != c and valid[q], (r, g, c, q)
            pair = (min(c, q), max(c, q))
            assert pair not in seen
            seen.add(pair)
            visitors.append(q)
            if g != 0 and (c - r == 1 or c - r == 1 - m):
                handing.append(g)
        assert len(set(visitors)) == ng and not set(visitors) & set(own)
        assert len(handing) <= 1
        for g in handing:                    # stored before the barrier
            valid[own[g]] = True
        for g in handing:                    # picked up after the barrier
            if r != m - 1:
                cn = r + 1 + h
                cn = cn - m if cn >= m else cn
                assert valid[cn], (r, g, cn)
                valid[cn] = False
                own[g] = cn
            else:
                own[g] = None
        if r != m - 1:
            for g in range(1, ng):
                assert 1 <= (own[g] - (r + 1)) % m <= h
    assert len(seen) == k * (k - 1) // 2


def test_wct_knobs_defaults_and_overrides(monkeypatch):
    """host-side knobs of the WCT container: eigensolver early stop by conv precision, eigenvalue truncation
    (util_wct.py:26-27), bounded CUDA-graph cache"""
    w = P.WCT(SimpleNamespace(mode="16x", numpy=False))
    old = P.get_precision()
    try:
        P.set_precision("tf32")
        assert w._early() == 1e-2
        w.dist = object()                     # sharded: tight setting regardless of the conv precision (tile invariance)
        assert w._early() == 1e-4
        w.dist = None
        P.set_precision("fp32")
        assert w._early() == 1e-4
        w.eig_early = 3e-6
        assert w._early() == 3e-6
    finally:
        P.set_precision(old)
    assert w._keep(128) == 0                                    # the shipped reference keeps every direction
    w.rat_eig = 0.25
    assert w._keep(128) == 32 and w._keep(24) == 6              # int(C * RatEigenValue)
    w.num_eig = 30
    assert w._keep(128) == 30                                   # NumEigenValue wins
    assert w.max_graphs >= 1 and len(w._graphs) == 0
    assert w.whiten_solver == "jacobi" and not w._use_ns(128)
    w.whiten_solver = "ns"
    assert not w._use_ns(128)                                   # truncation knobs need eigenvalues: stays on the eigensolver
    w.num_eig = w.rat_eig = None
    assert w._use_ns(128) and w._use_ns(24) and not w._use_ns(256)
    monkeypatch.setenv("WCTB_EIG_EARLY", "1e-3")
    assert P.WCT(SimpleNamespace(mode="16x", numpy=False))._early() == 1e-3


@pytest.mark.parametrize("solver,numpy_flag,alpha", [("jacobi", False, 1.0), ("jacobi", True, 0.6), ("ns", False, 1.0), ("ns", True, 0.6)])
def test_wct_params_plumbing_with_cpu_stand_ins(monkeypatch, solver, numpy_flag, alpha):
    """WCT._wct_params (statistics -> solver -> matrix) with the libwctb calls replaced by torch-cpu stand-ins that
    follow the documented contracts of include/wctb.h: checks the host-side composition (argument order, scales, +I on the
    content side only, both solver branches) against the oracle's transform on CPU.  The kernels themselves are covered
    by the GPU suite."""
    from collaborative_distillation_b200 import ops, util_wct

    def to_p4(x):                                   # [C,H,W] -> [C/4,H,W,4]
        C, H, W = x.shape
        return x.view(C // 4, 4, H, W).permute(0, 2, 3, 1).contiguous()

    def flat(x_p4, region):
        y0, y1, x0, x1 = region if region is not None else (0, x_p4.shape[1], 0, x_p4.shape[2])
        return x_p4[:, y0:y1, x0:x1, :].permute(0, 3, 1, 2).reshape(x_p4.shape[0] * 4, -1).double()

    def channel_sum(x, region=None):
        return flat(x, region).sum(1)

    def centered_gram(x, mean, region=None, out=None, fast=False):
        xc = flat(x, region) - mean[:, None]
        out += xc @ xc.t()
        return out

    def eigh(a, scale, add_identity=False, return_sweeps=False, early_stop=None):
        assert early_stop is not None
        ev, evec = [], []
        for p in range(a.shape[0]):
            S = a[p] * float(scale[p]) + (torch.eye(a.shape[1], dtype=torch.float64) if add_identity else 0)
            w, v = torch.linalg.eigh(S)
            ev.append(w.clamp_min(0))
            evec.append(v.t().contiguous())          # row k = eigenvector k (evecs[k*C + i])
        return torch.stack(ev), torch.stack(evec)

    def spectral(e, v, tau, power):
        keep = e > tau * e.max()
        return (v.t()[:, keep] * e[keep].pow(power)) @ v[keep]

    def finish(W, Col, c_mean, s_mean, alpha):
        C = len(c_mean)
        M = alpha * Col @ W + (1 - alpha) * torch.eye(C, dtype=torch.float64)
        return M.float(), (alpha * s_mean + (1 - alpha) * c_mean).float(), c_mean.float()

    def wct_matrix(c_e, c_v, c_mean, s_e, s_v, s_mean, tau, alpha, keep_c=0, keep_s=0):
        assert keep_c == 0 and keep_s == 0
        return finish(spectral(c_e, c_v, tau, -0.5), spectral(s_e, s_v, tau, 0.5), c_mean, s_mean, alpha)

    def whiten_ns(gram, scale, add_identity=False, return_info=False):
        S = gram * scale + (torch.eye(len(gram), dtype=torch.float64) if add_identity else 0)
        w, v = torch.linalg.eigh(S)
        keep = w > 1e-10 * w.max()
        return (v[:, keep] * w[keep].pow(-0.5)) @ v[:, keep].t()

    def wct_matrix_w(W, c_mean, s_e, s_v, s_mean, tau, alpha):
        return finish(W, spectral(s_e, s_v, tau, 0.5), c_mean, s_mean, alpha)

    for name, fn in (("channel_sum", channel_sum), ("centered_gram", centered_gram), ("eigh_jacobi", eigh), ("wct_matrix", wct_matrix),
                     ("whiten_ns", whiten_ns), ("wct_matrix_w", wct_matrix_w)):
        monkeypatch.setattr(ops, name, fn)
    g = torch.Generator().manual_seed(5)
    C = 24
    cF = torch.relu(torch.randn(C, C, generator=g) @ torch.randn(C, 20 * 30, generator=g) + 0.5).view(C, 20, 30)
    sF = torch.relu(torch.randn(C, C, generator=g) @ torch.randn(C, 16 * 18, generator=g) * 2 + 1.0).view(C, 16, 18)
    w = util_wct.WCT(SimpleNamespace(mode="16x", numpy=numpy_flag))
    w.whiten_solver = solver
    m, b, mc = w._wct_params(to_p4(cF), to_p4(sF), alpha)
    got = (m.double() @ (cF.view(C, -1).double() - mc.double()[:, None]) + b.double()[:, None]).view(1, C, 20, 30)
    ref = O.transform(cF, sF, alpha, numpy_variant=numpy_flag).double()
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
