#!/usr/bin/env python
"""Opcode- and line-level hotspots of one kernel from an ncu report's source page (needs -lineinfo / --import-source on).

    ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 | python tools/ncu_source_hotspots.py
"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
data = []
for r in rows[hi + 1:]:
    if len(r) == len(hdr) and r[iS].isdigit():
        data.append(r)
tot = sum(int(r[iS]) for r in data) or 1
toti = sum(int(r[iI]) for r in data) or 1
print(rows[0][:2])
print("total samples", tot, "total warp instructions", toti, "SASS lines", len(data))
ops, opi = collections.Counter(), collections.Counter()
for r in data:
    t = r[isrc].split()
    op = t[0] if t else "?"
    if op.startswith("@") and len(t) > 1:
        op = t[1]
    ops[op] += int(r[iS]); opi[op] += int(r[iI])
for op, c in ops.most_common(22):
    print("%-18s samples %6d (%5.1f%%)   instr %10d (%5.1f%%)" % (op, c, 100 * c / tot, opi[op], 100 * opi[op] / toti))
print("--- top lines")
for r in sorted(data, key=lambda r: -int(r[iS]))[:14]:
    print(r[iS], r[iI], r[isrc][:110])
