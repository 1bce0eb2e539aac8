#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total and
share of the GPU time.   usage: python scripts/launch_list_summary.py launches.csv out.md [skip_launches]"""
import collections
import csv
import sys

src, out = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3,
                  "second": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"].split("(")[0].replace("void ", ""), ms))
rows = rows[skip:]
tot = sum(ms for _, ms in rows)
agg = collections.OrderedDict()
for k, ms in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
lines = [f"# ncu launch list summary of `{src.split('/')[-1]}` ({len(rows)} launches, {tot:.3f} ms of kernel time; "
         "per-launch times are cold-cache and serialised)", "", "| kernel | launches | total ms | ms / launch | share |",
         "|---|---|---|---|---|"]
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {n} | {ms:.3f} | {ms / n:.3f} | {100 * ms / tot:.1f} % |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
