"""CPU tier: the pieces of bench.py that shape the ONE JSON line (no GPU, no timing): workloads of BASELINE.json,
roofline objects for a streaming kernel and for the latency-bound walker, the walker's CTA-lifetime spread."""
import json
import os

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workloads_cover_the_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert len(base["configs"]) == 5 and sorted(bench.WORKLOADS) == ["c1", "c2", "c3", "c4", "c5"]
    c3 = bench.workload_config("c3", 8)
    assert c3["scans_per_gpu"] == 256 and c3["yaw_candidates"] == 1024 and c3["points_per_scan"] == 120000
    assert c3["range_image"] == [112, 1440] and c3["parallelism"] == "scan-sharded x8" and "larger than L2" in c3["l2"]
    assert bench.workload_config("c2", 1)["range_image"] == [64, 2048] and bench.workload_config("c4", 1)["points_per_scan"] == 262144
    assert bench.WORKLOADS["c5"]["stream"] == 4541
    assert "NOT flushed" in bench.workload_config("c1", 1)["l2"]            # the single-scan case says so


class _FakeBench:
    n_scans = 256


def test_roofline_objects_have_the_contract_keys():
    streaming = {"ms_per_step": 0.62, "launches_per_step": 1.0, "share": 0.13, "bound": "hbm", "algorithmic_bytes_per_scan": 8730240,
                 "achieved_gbs": 3600.0, "frac": 0.55}
    r = bench.roofline_of("scatter_project", streaming, _FakeBench(), 6551.7, "measured")
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r
    assert r["frac"] <= 1.0 and r["bytes_per_launch"] == 8730240 * 256
    assert r["nominal_peak"] == 8000.0 and abs(r["frac_of_nominal"] - 3600.0 / 8000.0) < 1e-4      # both denominators
    walker = {"ms_per_step": 2.8, "launches_per_step": 1.0, "share": 0.6, "bound": "latency / issue (no streaming model)",
              "ncu": {"bytes_per_launch": 432000000, "source": "profiles/x.csv", "sm_throughput_pct": 15.6, "issue_active_pct": 28.8,
                      "warps_active_pct": 29.7}}
    r = bench.roofline_of("scan_walk", walker, _FakeBench(), 6551.7, "measured")
    assert r["traffic"] == 432000000 and 0 < r["frac"] < 0.1 and r["limiter"]["issue_active_pct"] == 28.8
    r = bench.roofline_of("scan_walk", {k: v for k, v in walker.items() if k != "ncu"}, _FakeBench(), 6551.7, "measured")
    assert r["achieved"] is None and r["frac"] is None and r["frac_of_nominal"] is None and "no ncu capture" in r["bytes_source"]


def test_walk_spread_percentiles():
    cyc = np.arange(1, 257, dtype=np.int64) * 19650            # 0.01 .. 2.56 ms at 1965 MHz
    tries = np.full(256, 11, dtype=np.int32); tries[-1] = 15
    s = bench.walk_spread({"walk_profile": (cyc, tries)}, {"sm_mhz": 1965.0})
    assert s["max"] == 2.56 and abs(s["p50"] - 1.29) < 0.02 and s["tries_max"] == 15 and s["tries_p50"] == 11


def test_committed_profile_records_are_consistent():
    """What bench.py reads from profiles/: the step traffic derived from the ncu launch list and the per-kernel records."""
    with open(os.path.join(ROOT, "profiles", "r2_step_traffic.json")) as f:
        t = json.load(f)
    assert t["dram_bytes_per_step"] == sum(k["dram_bytes"] for k in t["kernels"].values())
    assert abs(sum(k["share_of_device_time"] for k in t["kernels"].values()) - 1.0) < 0.01
    assert max(t["kernels"], key=lambda k: t["kernels"][k]["us"]) == "k_scan_walk"
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        n = json.load(f)
    for name in ("scan_walk", "scatter_project", "ingest_spherical", "close_fill_full", "compact_output"):
        assert n[name]["bytes_per_launch"] > 0 and n[name]["source"].startswith("profiles/r2_")
        assert os.path.exists(os.path.join(ROOT, n[name]["source"].split(" ")[0]))


def test_timed_mode_model_adds_walker_slot_time_and_streaming_kernels():
    """256 CTAs of 1.44 ms over 148 x 3 slots = 0.83 ms of device time next to 1.9 ms of streaming kernels."""
    table = {"scan_walk": {"ms_per_step": 2.9}, "scatter_project": {"ms_per_step": 0.62}, "ingest_spherical": {"ms_per_step": 0.48},
             "compact_output": {"ms_per_step": 0.30}, "close_fill_full": {"ms_per_step": 0.30}, "index_build": {"ms_per_step": 0.19}}
    cyc = {"total": int(1.44e-3 * 1965e6) * 256 * 10}
    m = bench.timed_mode_model("od", 256, table, cyc, 10, {"sm_mhz": 1965.0}, 2.95, sm_count=148)
    assert m["walker_cta_slots"] == 444
    assert abs(m["walker_slot_ms_per_step"] - 256 * 1.44 / 444) < 1e-3
    assert abs(m["streaming_kernels_ms_per_step"] - 1.89) < 1e-9
    assert abs(m["sum_ms"] - (m["walker_slot_ms_per_step"] + 1.89)) < 1e-3 and m["measured_ms_per_step"] == 2.95
    assert bench.timed_mode_model("ss", 256, table, cyc, 10, {}, 13.0, sm_count=148)["walker_cta_slots"] == 296
