#!/usr/bin/env python3
"""Summarise an ncu report: headline metrics + executed instructions / stall samples per source line
(all source files).  usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_ltcfabric.sum"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %s %s" % (w, vals[i], rows[1][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = []
fname = "?"; ci = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1] if len(r) > 1 else "?"; ci = None; continue
    if r[0] == "Line No":
        ci = {n: k for k, n in enumerate(r)}; continue
    if ci and r[0].isdigit() and "# Samples" in ci:
        try:
            agg.append((fname, int(r[0]), r[1][:100], int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["stall_long_sb"]] or 0)))
        except Exception:
            pass
ti = sum(a[4] for a in agg) or 1; ts = sum(a[3] for a in agg) or 1
print("total warp-instructions", ti, "samples", ts)
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%-18s %4d %6.2f%% inst %6.2f%% samp (long_sb %5.2f%%)  %s" % (a[0][:18], a[1], 100 * a[4] / ti, 100 * a[3] / ts, 100 * a[5] / ts, a[2]))
