#!/usr/bin/env python
"""Text summary of the decode ncu captures: launch lists (gpu__time_duration per launch, cold cache, serialised) and the
--set full reports of the step kernels.

    python tools/summarize_ncu_decode.py gpurun_out/r2_final > profiles/r02g_ncu_decode_summary.txt
"""
import collections
import csv
import re
import subprocess
import sys

prefix = sys.argv[1]


def launches(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for x in rows:
        name = re.sub(r"\(.*", "", x["Kernel Name"])[:56]
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1000 if x["Metric Unit"] == "ns" else (v * 1000 if x["Metric Unit"] == "ms" else v)
        a = agg.setdefault((name, x["Grid Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("== %s: %d launches, %.1f us in total (ncu: cold cache, serialised -- compare SHARES)" % (title, len(rows), tot))
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-58s grid %-16s n=%4d  total %9.1f us (%5.1f%%)  avg %8.2f us" % (name, grid, n, us, 100 * us / tot, us / n))
    print()


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration us"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read"),
            ("launch__registers_per_thread", "registers"), ("smsp__cycles_active.avg", "smsp cycles active"),
            ("sm__cycles_elapsed.avg", "sm cycles elapsed")]
    idx = [(hdr.index(k) if k in hdr else -1, lab) for k, lab in want]
    print("== %s (ncu --set full --clock-control none)" % title)
    for r in rows[2:]:
        print("  " + " | ".join("%s=%s%s" % (lab, r[i][:70].replace("void s2vt::xd::", "").replace("void s2vt::", ""),
                                            (" " + rows[1][i]) if rows[1][i] and lab not in ("kernel", "grid") else "") for i, lab in idx if i >= 0))
    print()


launches(prefix + "_ncu_greedy_launches.csv", "greedy, 512 videos (one call incl. weight preparation)")
launches(prefix + "_ncu_beam_launches.csv", "beam-5, 230 videos, depth 30 (one call incl. weight preparation)")
full(prefix + "_prof_greedy.ncu-rep", "greedy decode step kernels at 512 rows: word_rnn step (xgemm<64,1>) and vocab + argmax (xgemm<128,2>)")
full(prefix + "_prof_beam.ncu-rep", "beam step kernels at 1150 slots: LSTM steps (xgemm<128,1>), vocab + log-softmax + top-k (xgemm<128,3>), bookkeeping")
