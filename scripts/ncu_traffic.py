#!/usr/bin/env python
"""DRAM bytes per launch of every kernel captured in an .ncu-rep -> JSON fragment for profiles/ncu_traffic_r2.json.
usage: ncu_traffic.py <rep> <config> <source text> [existing json]  (read here, no GPU needed)"""
import csv
import io
import json
import subprocess
import sys

rep, config, source = sys.argv[1], sys.argv[2], sys.argv[3]
path = sys.argv[4] if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
    us = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    us *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1e-3)
    a = acc.setdefault(name, {"n": 0, "rd": 0.0, "wr": 0.0, "us": 0.0})
    a["n"] += 1
    a["rd"] += rd
    a["wr"] += wr
    a["us"] += us
frag = {k: {"dram_bytes": int((a["rd"] + a["wr"]) / a["n"]), "dram_read": int(a["rd"] / a["n"]),
            "dram_write": int(a["wr"] / a["n"]), "ncu_duration_us": round(a["us"] / a["n"], 3), "launches": a["n"],
            "source": source} for k, a in acc.items()}
doc = {}
if path:
    try:
        with open(path) as f:
            doc = json.load(f)
    except FileNotFoundError:
        pass
doc[config] = frag
text = json.dumps(doc, indent=1)
if path:
    with open(path, "w") as f:
        f.write(text + "\n")
else:
    print(text)
