#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) as a markdown table: per captured launch the duration,
DRAM traffic, DRAM / tensor-pipe utilisation, occupancy and registers.

    python scripts/ncu_summary.py gpurun_out/r01g_prof.ncu-rep > profiles/r01g_ncu_full.md
"""
import csv
import io
import subprocess
import sys

COLS = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1/smem %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs")]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(c) if c in hdr else None, n) for c, n in COLS]
    print("| " + " | ".join(n for _, n in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in rows[2:]:
        out = []
        for i, n in idx:
            v = r[i] if i is not None else ""
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "")[:40]
            else:
                try:
                    v = "%.1f" % float(v.replace(",", ""))
                except ValueError:
                    pass
            out.append(v)
        print("| " + " | ".join(out) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
