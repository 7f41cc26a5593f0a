"""Host-side bookkeeping of bench.py that the reported numbers rest on: the algorithmic FLOP count of a step, the
workload table (BASELINE.json configs), the bounded CPU sample of a workload and the per-N default configuration."""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bench = importlib.import_module("bench")


def test_algorithmic_flops_of_the_baseline_configs():
    # the figures DESIGN.md section 3 and the bench line quote (TFLOP per step)
    for cfg, want in (("cfg3", 2.0676), ("cfg4", 9.256), ("cfg5", 31.855)):
        (Hc, Wc), (Hs, Ws) = bench.CONFIGS[cfg]
        mode = "original" if cfg == "cfg5" else "16x"
        got = bench.algorithmic_conv_flops(mode, Hc, Wc, Hs, Ws) / 1e12
        assert abs(got - want) <= 5e-4 * want, (cfg, got)


def test_flops_count_every_conv_once():
    # stage 1 of the 16x nets alone: conv0 (1x1, 3->3) + 3->24 on content and style, 24->3 on content
    from collaborative_distillation_b200 import arch
    H = W = 64
    enc = sum(2 * 9 * L["cin"] * L["cout"] * H * W for L in arch.encoder_layers("16x", 1))
    dec = sum(2 * 9 * L["cin"] * L["cout"] * H * W for L in arch.decoder_layers("16x", 1))
    assert enc == 2 * 9 * 3 * 24 * H * W and dec == 2 * 9 * 24 * 3 * H * W
    total = bench.algorithmic_conv_flops("16x", H, W, H, W)
    assert total > 2 * (enc + 2 * 3 * 3 * H * W) + dec          # all five stages are in, stage 1 is the smallest part


def test_workloads_and_defaults():
    assert bench.default_config(1) == "cfg3" and bench.default_config(8) == "weak"
    Hc, Wc, Hs, Ws, name = bench.workload("cfg3", 1)
    assert (Hc, Wc, Hs, Ws) == (2160, 3840, 2000, 2000) and "3840x2160" in name and "--mode 16x" in name
    Hc, Wc, Hs, Ws, name = bench.workload("weak", 8)
    assert (Hc, Wc) == (2160, 3840 * 8) and "weak" in name
    assert "original" in bench.workload("cfg5", 2)[4]


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_cpu_sample_is_a_bounded_crop_with_the_same_proportions(cfg):
    (Hc, Wc), (Hs, Ws) = bench.CONFIGS[cfg]
    hc, wc, hs, ws = bench.cpu_sample_shape(Hc, Wc, Hs, Ws)
    assert hc <= Hc and wc <= Wc and hs <= Hs and ws <= Ws
    assert all(v % 16 == 0 for v in (hc, wc, hs, ws))
    assert hc * wc <= 0.62e6                                     # about half a megapixel of content per timed CPU step
    if hc < Hc:                                                  # aspect ratio and content : style area ratio are kept (to 16-px rounding)
        assert abs(wc / hc - Wc / Hc) <= 0.1 * Wc / Hc
        assert abs((hc * wc) / (hs * ws) - (Hc * Wc) / (Hs * Ws)) <= 0.15 * (Hc * Wc) / (Hs * Ws)


def test_roofline_helper_handles_a_pass_without_recorded_launches():
    assert bench.rows_available({}) is False
    assert bench.rows_available({("k", "s"): []}) is False
    assert bench.rows_available({("k", "s"): [1]}) is True
