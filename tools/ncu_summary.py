#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: one row per captured launch with the numbers the roofline
discussion uses (duration, DRAM bytes, pipe utilisation, issue rate, registers).  Runs where ncu is installed
(the build container reads reports brought back from the GPU box):
    python tools/ncu_summary.py gpurun_out/<run>/<name>.ncu-rep > profiles/<round>_ncu_<name>.md"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("launch__registers_per_thread", "regs"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__inst_executed.sum", "warp instr"),
]


def short(name):
    name = name.replace("void ", "").replace("tsc::", "")
    return name.split("(")[0]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch numbers, cold-cache and serialised)\n")
    print("| # | kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for r in body:
        cells = []
        for key, _ in COLS:
            if key not in idx:
                cells.append("-")
                continue
            v, u = r[idx[key]], units[idx[key]]
            try:
                f = float(v)
                if key.startswith("dram__bytes"):
                    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    cells.append(f"{f * scale:.2f} MB")
                elif key == "gpu__time_duration.sum":
                    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3,
                             "s": 1e6, "second": 1e6}.get(u, 1.0)
                    cells.append(f"{f * scale:.1f}")
                elif key == "smsp__inst_executed.sum":
                    cells.append(f"{f / 1e6:.1f} M")
                elif key == "launch__registers_per_thread":
                    cells.append(f"{int(f)}")
                else:
                    cells.append(f"{f:.1f}")
            except ValueError:
                cells.append(v)
        grid = r[idx["Grid Size"]].replace(", 1, 1", "").strip("()")
        block = r[idx["Block Size"]].replace(", 1, 1", "").strip("()")
        print(f"| {r[idx['ID']]} | `{short(r[idx['Kernel Name']])}` | {grid} x {block} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
