#!/usr/bin/env python
"""Summarise ncu outputs for profiles/ (runs in the build container, no GPU needed).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv                > profiles/rNN_ncu_launch_list_summary.txt
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep [traffic.json]     > profiles/rNN_ncu_full_top_kernels.txt

`launches`: the CSV written by `ncu --metrics gpu__time_duration.sum --csv --log-file`.  `full`: a `--set full` report; prints the
per-launch headline metrics and (optionally) writes per-kernel-family DRAM traffic as JSON for bench.py's roofline.traffic.
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    return re.sub(r"<.*", "", name)


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
        k = short(r[ik])
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print("%-46s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-46s %6d %12.1f %6.1f%%" % (k, c, t, 100 * t / tot))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(rep, traffic_out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    traffic = {}
    for r in rows[2:]:
        vals = []
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals.append("%s=%s %s" % (k.split(".")[0], r[i], units[i]))
                d[k] = (r[i], units[i])
        print("k=%-34s | " % short(r[ik])[:34] + " | ".join(vals))
        try:
            def to_bytes(key):
                v, u = d[key]
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            fam = short(r[ik])
            c, b = traffic.get(fam, (0, 0.0))
            traffic[fam] = (c + 1, b + to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
        except Exception:
            pass
    if traffic_out:
        json.dump({k: {"launches": c, "dram_bytes_per_launch": b / c} for k, (c, b) in traffic.items()}, open(traffic_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
