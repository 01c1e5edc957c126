"""Host-side pieces of bench.py that need no GPU: the sha gate of `roofline.traffic`, the workload description shared by
both arms, the (QP, frame) shard arithmetic bench.py and Inference_QBD share."""
import json
import os

import bench
from pmp_vvc_tip2023_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_summary_matches_committed_kernel_sources():
    """The committed ncu DRAM summary was measured on exactly the kernel sources in the tree (bench.py reports
    `roofline.traffic` only then)."""
    traffic, src = bench.measured_traffic(4800)
    assert traffic is not None, src
    tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    assert abs(traffic - tj["conv_tc_dram_bytes_per_launch_per_block"] * 4800) < 1.0
    # per launch and block the conv kernels move about the compulsory bytes (0.3 .. 1 MB), not multiples of them
    assert 0.3e6 < tj["conv_tc_dram_bytes_per_launch_per_block"] < 1.0e6


def test_stale_traffic_summary_is_refused(monkeypatch):
    monkeypatch.setattr(bench, "kernel_source_sha", lambda: "0" * 16)
    traffic, why = bench.measured_traffic(4800)
    assert traffic is None and "not reported" in why


def test_both_arms_carry_the_same_workload_config():
    a, b = bench.workload_config(), bench.workload_config("tc")
    assert a == b and a["units_per_step"] == 4800 and "configs[1]" in a["workload"]


def test_bench_shards_tile_the_4k_sequence():
    for world in (1, 2, 4, 8):
        per_qp = {}
        for r in range(world):
            for qi, a, b in sharding.qp_frame_shards(30, 4, world, r):
                per_qp.setdefault(qi, []).append((a, b))
        assert sorted(per_qp) == [0, 1, 2, 3]
        assert all(s[0][0] == 0 and s[-1][1] == 30 for s in per_qp.values())
