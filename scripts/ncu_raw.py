"""Print the headline raw metrics of each kernel launch in an ncu report.  usage: python scripts/ncu_raw.py report.ncu-rep [kernel-regex]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum"]
STALL = "smsp__average_warps_issue_stalled_"
cmd = ["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    cmd += ["-k", "regex:" + sys.argv[2]]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print("==", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k:90s} {d[k]}")
    st = {k[len(STALL):-len('_per_issue_active.ratio')]: float(v) for k, v in d.items()
          if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
    print("  stalls/issue:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
