"""Summarise an ncu report per kernel from the SASS source page: stall mix and hottest instructions.
usage: python scripts/ncu_hot.py <report.ncu-rep> <kernel-regex> [top]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
H = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[H["# Samples"]]) for r in data)
print("kernel", kre, "instructions", len(data), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
mix = {s: sum(int(r[H[s]]) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]}={100*v/max(1,tot):.1f}%" for k, v in sorted(mix.items(), key=lambda x: -x[1]) if v))
exc = sum(int(r[H["L1 Wavefronts Shared Excessive"]]) for r in data)
wf = sum(int(r[H["L1 Wavefronts Shared"]]) for r in data)
print(f"shared wavefronts {wf} excessive {exc} ({100*exc/max(1,wf):.1f}%)")
ie = sum(int(r[H["Instructions Executed"]]) for r in data)
print("warp instructions executed", ie)
order = sorted(range(len(data)), key=lambda i: -int(data[i][H["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    st = {s[6:]: int(r[H[s]]) for s in stalls if int(r[H[s]])}
    main = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{i:5d} {int(r[H['# Samples']]):7d} ({100*int(r[H['# Samples']])/tot:4.1f}%) exec={r[H['Instructions Executed']]:>9s} "
          f"shx={r[H['L1 Wavefronts Shared Excessive']]:>9s} {r[H['Source']].strip()[:70]:70s} {main}")
