#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU needed): one row per captured launch."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"),
        ("dram__bytes_write.sum", "wrMB"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit"), ("lts__t_sector_hit_rate.pct", "L2hit"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__grid_size", "grid"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("smsp__inst_executed.sum", "inst")]
idx = {n: hdr.index(n) for n, _ in want if n in hdr}
units = rows[1]
print("| " + " | ".join(l for n, l in want if n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for n, l in want:
        if n not in idx:
            continue
        v = r[idx[n]]
        u = units[idx[n]]
        if l == "kernel":
            v = v.split("(")[0][:44]
        elif l == "us":
            x = float(v.replace(",", ""))
            v = "%.1f" % (x / 1000 if u in ("ns", "nsecond") else x if u.startswith("us") else x * 1000 if u.startswith("ms") else x)
        elif l in ("rdMB", "wrMB"):
            x = float(v.replace(",", ""))
            f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            v = "%.1f" % (x * f)
        else:
            try:
                v = "%.1f" % float(v.replace(",", ""))
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
