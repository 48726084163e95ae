"""bench.py bookkeeping that the roofline numbers rest on (CPU only; no GPU work, no oracle timing)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import bench_denoise  # noqa: E402


def test_reproj_algorithmic_bytes_matches_survey():
    """16 B per point-view + per view 2 x 6 x 512^2 x 8 B z-buffer + 1000 x 2000 x 3 B output;
    N = 2.54 M, V = 24 -> 1.72 GB per set -> 0.263 ms at 6551.7 GB/s."""
    c = bench.REPROJ_CFG
    b = bench.reproj_algorithmic_bytes(2_538_200, c["V"], c["face_res"], c["pano"])
    assert b == 24 * (16 * 2_538_200 + 2 * 6 * 512 * 512 * 8 + 1000 * 2000 * 3)
    assert abs(b / 1e9 - 1.7226) < 1e-3
    assert abs(b / 6551.7e9 * 1e3 - 0.263) < 1e-3


def test_denoise_algorithmic_flops_match_survey_appendix_c():
    from evoworld_b200.unet import DEFAULT_CONFIG, algorithmic_flops

    cfg = dict(DEFAULT_CONFIG, **bench_denoise.UNET_CFG)
    f14 = algorithmic_flops(cfg, 2, 14, 72, 128)
    assert abs(f14["total"] - 89.62) < 0.05 and abs(f14["sdpa_spatial"] - 17.37) < 0.02
    assert abs(algorithmic_flops(cfg, 2, 25, 72, 128)["total"] - 159.9) < 0.3
    assert abs(algorithmic_flops(cfg, 2, 14, 128, 256)["total"] - 476.4) < 0.5


def test_ncu_traffic_file_is_consistent():
    d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    r = d["reproj_24view_set"]
    assert r["bytes"] == r["passes"] * sum(r["per_pass_bytes"].values())
    assert bench.ncu_traffic("reproj_24view_set") == r["bytes"]
    assert bench.ncu_traffic("no such kernel") is None
    for src in r["source"].split(" ")[:1]:
        assert os.path.exists(os.path.join(ROOT, src))


def test_measured_peaks_fallback_and_file():
    p = bench.measured_peaks()
    assert 5000 < p["hbm_gbs"] < 8000 and 1000 < p["tf_sustained"] <= p["tf_burst"] < 2300
    assert "measured" in p["source"] or "fallback" in p["source"]
