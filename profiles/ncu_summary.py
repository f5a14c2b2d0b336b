#!/usr/bin/env python
"""Per-launch summary of an ncu report: python profiles/ncu_summary.py X.ncu-rep  (runs `ncu -i ... --page raw --csv`)"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
H, U = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
idx = [(H.index(a), b) for a, b in want if a in H]
for r in rows[2:]:
    parts = []
    for i, b in idx:
        v = r[i]
        if b == "kernel":
            v = v.replace("void ", "").replace("cone::", "").replace("unnamed>::", "").replace("<unnamed>::", "")[:34]
        else:
            try:
                v = f"{float(v):.2f}"
            except ValueError:
                pass
            v = f"{b}={v}{U[i] if b in ('rd','wr') else ''}"
        parts.append(v)
    print(" ".join(parts))
