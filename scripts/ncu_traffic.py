"""DRAM bytes per kernel from an `ncu --set full` report -> profiles/r1_ncu_traffic.json (read by bench.py for
`roofline.traffic`).  usage: python scripts/ncu_traffic.py report.ncu-rep out.json [keep.json]
Kernels missing from the report (captured in a separate metrics-only pass) are carried over from keep.json."""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    keep = json.load(open(sys.argv[3]))["kernels"] if len(sys.argv) > 3 else {}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ni = hdr.index("Kernel Name")
    ri, wi, ti = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    kernels = dict(keep)
    for r in data:
        name = r[ni].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
        tms = float(r[ti]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[ti], 1.0)
        try:
            rd, wr = float(r[ri]) * UNIT[units[ri]], float(r[wi]) * UNIT[units[wi]]
        except ValueError:
            continue
        if rd != rd or wr != wr:        # NaN: metric not collected for this launch, keep the earlier pass
            continue
        kernels[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                         "time_ms_under_ncu": tms}
    json.dump({"source": f"{rep.split('/')[-1]} (ncu --set full --clock-control none, bench.py --steps 1 --warmup 3, one launch "
                         "per kernel over the whole config-2 manifest); kernels absent from that capture carried over from "
                         "the previous metrics-only pass", "kernels": kernels}, open(out, "w"), indent=1)
    for k, v in kernels.items():
        print(f"{k:24s} read {v['dram_read_bytes'] / 1e9:8.3f} GB  write {v['dram_write_bytes'] / 1e9:8.3f} GB  {v['time_ms_under_ncu']:8.3f} ms")


if __name__ == "__main__":
    main()
