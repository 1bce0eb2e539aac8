"""Segment a kernel's SASS (ncu source page) into runs of equal execution count: per run the warp-instruction
total, its share, the samples and an opcode histogram.  usage: python scripts/ncu_segments.py report.ncu-rep kernel-regex [launch_index]"""
import collections
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
launches, hdr = [], None
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        launches.append([])
        continue
    if hdr and len(r) == len(hdr):
        launches[-1].append(r)
data = launches[which]
H = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[H["Instructions Executed"]]) for r in data)
tots = sum(int(r[H["# Samples"]]) for r in data)
print(f"launch {which}: {len(data)} SASS instructions, {tot} warp instructions, {tots} samples")
seg = []
for i, r in enumerate(data):
    e = int(r[H["Instructions Executed"]])
    if seg and (seg[-1]["e"] == e or (e and abs(seg[-1]["e"] - e) / max(e, seg[-1]["e"]) < 0.02)):
        s = seg[-1]
    else:
        s = {"e": e, "start": i, "n": 0, "sum": 0, "smp": 0, "ops": collections.Counter(), "bar": 0}
        seg.append(s)
    s["n"] += 1
    s["sum"] += e
    s["smp"] += int(r[H["# Samples"]])
    s["bar"] += int(r[H["stall_barrier"]]) if "stall_barrier" in H else 0
    op = r[H["Source"]].strip().split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
    s["ops"][op.split(".")[0]] += 1
for s in seg:
    if s["sum"] / tot < 0.004 and s["smp"] / tots < 0.004:
        continue
    ops = ",".join(f"{k}:{v}" for k, v in s["ops"].most_common(7))
    print(f"[{s['start']:5d}+{s['n']:4d}] exec/inst {s['e']:>11d}  inst-share {100*s['sum']/tot:5.1f}%  samples {100*s['smp']/tots:5.1f}%  barrier-smp {100*s['bar']/tots:4.1f}%  {ops}")
