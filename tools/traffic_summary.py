"""Summarise an ncu per-launch DRAM-byte capture into profiles/rNN_traffic.json (what bench.py's roofline.traffic reads).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/dram.csv python tools/profile_run.py 2 480        (GPU box: 2 frames = 960 blocks, 480-block chunks)
    python tools/traffic_summary.py gpurun_out/dram.csv profiles/r02_traffic.json 960 480

The JSON records the sha of the kernel sources it was measured on (bench.kernel_source_sha): bench.py refuses to report
`traffic` from a file whose sha differs from the loaded sources.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, dst, blocks, chunk = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    import bench
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"kernel": r[ik]})
        d[r[im]] = float(r[iv].replace(",", ""))
    # the workload runs predict_frames twice (warm-up + measured): keep the second half of the launches
    ids = list(per)
    half = ids[len(ids) // 2:]
    cls = collections.defaultdict(lambda: {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0})
    for i in half:
        d = per[i]
        name = d["kernel"]
        key = "conv_tc" if "conv_tc" in name else ("pool2_split" if "pool2" in name else name.split("(")[0].split("::")[-1].split("<")[0])
        c = cls[key]
        c["launches"] += 1
        c["read"] += d.get("dram__bytes_read.sum", 0.0)
        c["write"] += d.get("dram__bytes_write.sum", 0.0)
        c["ns"] += d.get("gpu__time_duration.sum", 0.0)
    tc = cls["conv_tc"]
    total = sum(c["read"] + c["write"] for c in cls.values())
    launches_per_chunk = tc["launches"] / max(1, blocks // chunk)
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (tools/profile_run.py: %d blocks, %d-block chunks), %s"
                     % (blocks, chunk, os.path.basename(src)),
           "kernel_source_sha": bench.kernel_source_sha(),
           "conv_tc_dram_bytes_per_launch_per_block": (tc["read"] + tc["write"]) / max(tc["launches"], 1) / chunk,
           "conv_tc_launches": tc["launches"], "conv_tc_launches_per_chunk": launches_per_chunk,
           "conv_tc_read_bytes": tc["read"], "conv_tc_write_bytes": tc["write"],
           "all_kernels_dram_bytes_per_block_pair": total / blocks,
           "classes": {k: {"launches": c["launches"], "dram_bytes_per_block_pair": (c["read"] + c["write"]) / blocks,
                           "ncu_time_ms": c["ns"] * 1e-6} for k, c in cls.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
