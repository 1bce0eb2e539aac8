"""Histogram of an ncu source page by SASS region: share of samples / executed instructions.
usage: python scripts/ncu_regions.py <report.ncu-rep> <kernel-regex> [region_size]"""
import collections
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
H = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[H["# Samples"]]) for r in data)
totex = sum(int(r[H["Instructions Executed"]]) for r in data)
print("instructions", len(data), "samples", tot, "warp-insts", totex)
for lo in range(0, len(data), step):
    seg = data[lo:lo + step]
    smp = sum(int(r[H["# Samples"]]) for r in seg)
    ex = sum(int(r[H["Instructions Executed"]]) for r in seg)
    mx = max(int(r[H["Instructions Executed"]]) for r in seg)
    ops = collections.Counter((r[H["Source"]].split()[1] if r[H["Source"]].strip().startswith('@')
                               else r[H["Source"]].split()[0]).split('.')[0] for r in seg)
    print(f"{lo:5d} {100 * smp / tot:5.1f}% smp {100 * ex / totex:5.1f}% inst maxexec {mx:9d}", ops.most_common(5))
