#!/usr/bin/env python
"""Top instructions by warp-stall samples of one kernel of an .ncu-rep (SASS view, read here, no GPU needed).
usage: ncu_hot_sass.py <rep> <kernel regex> [N]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[h]
iS, iSrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
iE = hdr.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[h + 1:]):
    if len(r) <= iS or not r[iS].strip().isdigit():
        continue
    data.append((int(r[iS]), idx, r[iSrc], r[iE]))
tot = sum(d[0] for d in data)
print("kernel", kern, "total samples", tot, "instructions", len(data))
for s, idx, src, ex in sorted(sorted(data, reverse=True)[:N], key=lambda x: x[1]):
    print("%5d %6d %5.1f%%  exec %-9s %s" % (idx, s, 100.0 * s / max(tot, 1), ex, src[:100]))
