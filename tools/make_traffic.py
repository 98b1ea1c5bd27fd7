#!/usr/bin/env python3
"""profiles/traffic.json from full ncu captures: DRAM bytes (read + write) per launch of a hot kernel,
read by bench.py for roofline.traffic.  usage: make_traffic.py kernel=report.ncu-rep[:summary-path] ... [--size 512 --order 4]"""
import csv, io, json, subprocess, sys
from pathlib import Path
out = Path(__file__).resolve().parent.parent / "profiles" / "traffic.json"
size, order = 512, 4
args = [a for a in sys.argv[1:]]
res = json.loads(out.read_text()) if out.exists() else {}
for a in args:
    if a.startswith("--size="):
        size = int(a.split("=")[1]); continue
    if a.startswith("--order="):
        order = int(a.split("=")[1]); continue
    name, rep = a.split("=", 1)
    src = rep
    if ":" in rep:
        rep, src = rep.split(":", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    def get(metric):
        i = hdr.index(metric)
        v = float(vals[i]); u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[u]
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    res[name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "size": size, "order": order, "source": src}
out.write_text(json.dumps(res, indent=1) + "\n")
print(out.read_text())
