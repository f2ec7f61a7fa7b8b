"""Provenance of the ncu figures bench.py quotes in `roofline.binding` / `roofline.traffic`: profiles/extend_ncu_metrics.json
must be what tools/ncu_metrics.py derives from the committed machine-readable export of the capture (the raw page of
the .ncu-rep as CSV), not hand-copied numbers.  CPU only."""
import csv
import json
import os

import uvrt_testlib as T

PROFILES = os.path.join(T.ROOT, "profiles")


def _num(v):
    return float(v.replace(",", ""))


def test_bench_binding_figures_come_from_the_committed_ncu_export():
    m = json.load(open(os.path.join(PROFILES, "extend_ncu_metrics.json")))
    raw = os.path.join(PROFILES, m["raw_csv"])
    assert os.path.exists(raw), "the CSV export named in extend_ncu_metrics.json is not committed"
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if "k_extend_fast" in dict(zip(hdr, r)).get("Kernel Name", "")]
    assert body, "no k_extend_fast instance in the export"
    d, u = dict(zip(hdr, body[0])), dict(zip(hdr, units))
    assert d["Kernel Name"] == m["kernel"]
    pairs = {
        "l1tex_data_pipe_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
        "warp_instructions": "smsp__inst_executed.sum",
        "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
        "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "long_scoreboard_stalls_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "registers": "launch__registers_per_thread",
    }
    for key, metric in pairs.items():
        assert abs(m[key] - _num(d[metric])) <= 1e-9 * max(1.0, abs(m[key])), key
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = _num(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = _num(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    assert abs(m["bytes_per_launch"] - (rd + wr)) <= 1e-6 * (rd + wr)
    # the shipped kernel: 46 registers, no spills at 40 resident warps per SM
    assert m["registers"] <= 48


def test_traversal_statistics_behind_the_roofline_are_committed():
    st = json.load(open(os.path.join(PROFILES, "traversal_stats_route.json")))
    # B_ray = 44 + 64 I + 52 T; the bench multiplies it with the measured launch rate
    assert 20 < st["inner_visits_per_ray"] < 40 and 1 < st["tri_tests_per_ray"] < 5
