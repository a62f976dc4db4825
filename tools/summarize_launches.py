#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python tools/summarize_launches.py file.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    if "at::native" in name or "void at" in r["Kernel Name"]:
        m = re.search(r"(\w+_kernel\w*|\w+Functor\w*|\w+_impl\w*)", r["Kernel Name"])
        name = "torch:" + (m.group(1) if m else name[-40:])
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.3f} ms total")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.3f} ms {100 * v[1] / tot:5.1f}%  x{v[0]:4d}  {k}")
