#!/usr/bin/env python3
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel count / median / min / max (us)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    d[re.sub(r"\(.*", "", r[ki])[:64]].append(v)
for k, v in d.items():
    v2 = sorted(v)
    print(f"{k:64s} n={len(v):4d} med={v2[len(v2) // 2] / 1000:8.1f}us  min={v2[0] / 1000:7.1f} max={v2[-1] / 1000:7.1f}")
