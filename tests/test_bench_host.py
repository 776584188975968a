"""Host-side pieces of bench.py that need no GPU: workload table, roofline denominators, gather policy."""

import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_workloads_cover_the_baseline_configs():
    b = _bench()
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5
    for name, (n, d, k) in {"c1": (1024, 2, 100), "c2": (65536, 128, 500), "c3": (65536, 784, 20), "c4": (262144, 64, 20),
                            "c5": (65536, 784, 20)}.items():
        assert b.WORKLOADS[name][1:] == (n, d, k)
    assert "c5" in b.WEAK and b.METRIC == "langevin_chain_steps_per_sec"
    # the north_star MLP target and its three-hidden-layer sibling: same chains, width and burst length
    assert b.WORKLOADS["mlp128"][1:] == b.WORKLOADS["mlp128x3"][1:] == (65536, 128, 100)


def test_tensor_roofline_counts_every_product_of_the_energy():
    """4 * (D*H1 + H1*H2 + H2) FLOP per chain-step for two hidden layers; a third adds 4 * H2*H3."""
    b = _bench()
    peaks = {"hbm_gbs": 6539.5, "bf16_tflops": 1665.4}
    units = 65536 * 100
    two = b.roofline("mlp128", peaks, "measured", 0.0, None, 0, 1.0, units)
    three = b.roofline("mlp128x3", peaks, "measured", 0.0, None, 0, 1.0, units)
    assert two["bound"] == three["bound"] == "tensor"
    assert two["algorithmic_flops_per_launch"] == 4 * (128 * 128 * 2 + 128) * units
    assert three["algorithmic_flops_per_launch"] == 4 * (128 * 128 * 3 + 128) * units
    wide = b.roofline("c3", peaks, "measured", 0.0, None, 0, 1.0, 65536 * 20)
    assert wide["algorithmic_flops_per_launch"] == 4 * (784 * 128 + 128 * 128 + 128) * 65536 * 20


def test_measured_peaks_fallback_and_partial_file(tmp_path, monkeypatch):
    b = _bench()
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    peaks, kind = b.measured_peaks()
    assert kind == "fallback" and peaks == {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6437.3, "bf16_tflops": 1675.2, "bf16_tflops_sustained": 1409.4}))
    peaks, kind = b.measured_peaks()
    assert kind == "measured" and peaks["hbm_gbs"] == 6437.3 and peaks["bf16_tflops"] == 1675.2
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6437.3}))
    peaks, kind = b.measured_peaks()
    assert kind.startswith("mixed") and peaks["bf16_tflops"] == 1590.0
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert b.measured_peaks()[1] == "fallback"


def test_c5_gather_policy():
    b = _bench()
    assert [b.c5_gather_mode(w) for w in (2, 4, 8)] == ["fused"] * 3   # peer stores from inside the burst kernel
    b.C5_GATHER = "nccl"
    assert b.c5_gather_mode(2) == "nccl"
