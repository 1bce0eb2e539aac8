#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small tracked CSV/markdown under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_v2_ncu_full.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe_fp64_%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "pipe_alu_%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefronts_%"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_%"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    lines = [f"# ncu --set full summary of `{rep.split('/')[-1]}` (one launch per kernel, --clock-control none)", ""]
    names = [r[name_i].split("(")[0].replace("void ", "") for r in data]
    lines.append("| metric | unit | " + " | ".join(names) + " |")
    lines.append("|---|---|" + "---|" * len(names))
    seen = set()
    for key, label in KEYS:
        if key not in hdr or label in seen:
            continue
        seen.add(label)
        i = hdr.index(key)
        vals = []
        for r in data:
            try:
                v = float(r[i])
                vals.append(f"{v:.4g}" if abs(v) < 1e6 else f"{v:.4e}")
            except ValueError:
                vals.append(r[i])
        lines.append(f"| {label} (`{key}`) | {units[i]} | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
