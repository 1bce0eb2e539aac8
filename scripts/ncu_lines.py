"""Per CUDA source line: warp instructions executed and stall samples (ncu --import-source on report).
usage: python scripts/ncu_lines.py report.ncu-rep kernel-regex [min_share_pct]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
mn = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, agg, seen_fn = None, None, {}, 0
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Function Name":
        seen_fn += 1
        continue
    if r and r[0] == "Line No":
        hdr = r
        H = {}
        for i, h in enumerate(hdr):
            H.setdefault(h, i)
        continue
    if not hdr or len(r) != len(hdr) or seen_fn > len(set([fname])) * 1000:
        continue
    if r[H["Line No"]] and r[H["Line No"]].isdigit() and r[H["Address"]] in ("", "-"):
        key = (fname, int(r[H["Line No"]]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        try:
            a[0] += int(r[H["Instructions Executed"]] or 0)
            a[1] += int(r[H["# Samples"]] or 0)
            a[2] += int(r[H["stall_barrier"]] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions (all launches in report) {tot}, samples {tots}")
for (f, ln), a in sorted(agg.items()):
    if 100 * a[0] / tot >= mn or 100 * a[1] / tots >= mn:
        print(f"{f}:{ln:4d}  inst {100*a[0]/tot:5.2f}%  samples {100*a[1]/tots:5.2f}%  barrier {100*a[2]/tots:5.2f}%  {a[3][:100]}")
