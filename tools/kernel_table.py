#!/usr/bin/env python
"""Per-kernel table of one bench step from an ncu metrics pass that recorded, for every launch,
gpu__time_duration.sum, dram__bytes_read.sum and dram__bytes_write.sum:

    python tools/kernel_table.py profiles/r1g_step_time_dram.csv [--traffic-json profiles/traffic.json] [--hbm-peak 6544.7]

Prints a markdown table (launches, ms per step, share, DRAM GB per step, achieved DRAM GB/s, % of the measured HBM peak) and,
with --traffic-json, rewrites the '<kernel>' entries (average DRAM bytes per launch) that bench.py reports as roofline.traffic.
The times are serialised, cold-cache ncu replays: shares and bytes are the evidence, not the absolute times."""
import argparse
import csv
import json
import re
from collections import defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--traffic-json")
ap.add_argument("--hbm-peak", type=float, default=6544.7)
ap.add_argument("--step-marker", default=None,
                help="a kernel launched exactly once at the start of every step (pack_batched_kernel): keep only the launches of the LAST "
                     "complete step, i.e. from the last-but-one marker launch up to the last one")
a = ap.parse_args()

per_launch = defaultdict(dict)
names = {}
for r in csv.DictReader(l for l in open(a.csv) if l.startswith('"')):
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per_launch[r["ID"]][r["Metric Name"]] = v * scale
    names[r["ID"]] = r["Kernel Name"]


def short(name):
    if "at::" in name:
        m = re.search(r"(\w+_kernel\w*)", name)
        return "torch:" + (m.group(1) if m else name[:40])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"[<(].*", "", name)
    return name.replace("b200::", "")


if a.step_marker:
    ids = sorted(per_launch, key=int)
    marks = [i for i in ids if a.step_marker in names[i]]
    if len(marks) >= 2:
        lo, hi = int(marks[-2]), int(marks[-1])
        per_launch = {i: m for i, m in per_launch.items() if lo <= int(i) < hi}
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per_launch.items():
    k = short(names[i])
    agg[k][0] += 1
    agg[k][1] += m.get("gpu__time_duration.sum", 0.0)
    agg[k][2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
print(f"{len(per_launch)} launches, {tot:.3f} ms serialised\n")
print("| kernel | launches | ms / step | share | DRAM GB / step | DRAM GB/s | % of HBM peak |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if ms < 0.02:
        continue
    gbs = by / ms / 1e6 if ms else 0.0
    print(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % | {by / 1e9:.3f} | {gbs:.0f} | {100 * gbs / a.hbm_peak:.0f} % |")
if a.traffic_json:
    try:
        tab = json.load(open(a.traffic_json))
    except FileNotFoundError:
        tab = {}
    for k, (n, ms, by) in agg.items():
        if not k.startswith("torch:"):
            tab[k] = int(by / n)
    tab["_source"] = a.csv
    json.dump(tab, open(a.traffic_json, "w"), indent=1)
