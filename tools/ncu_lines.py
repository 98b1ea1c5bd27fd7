#!/usr/bin/env python3
"""Summarise an ncu report: headline metrics + executed instructions / stall samples per source line.
usage: ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print("%-70s %s %s" % (w, vals[i], rows[1][i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
tabs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
t = tabs[0]
h = rows[t]
ci = {n: k for k, n in enumerate(h)}
agg = []
i = t + 1
while i < len(rows) and rows[i] and rows[i][0] != "File Path":
    r = rows[i]
    try:
        agg.append((int(r[0]), r[1][:100], int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["stall_long_sb"]] or 0)))
    except Exception:
        pass
    i += 1
ti = sum(a[3] for a in agg) or 1; ts = sum(a[2] for a in agg) or 1
print("total warp-instructions (first file)", ti, "samples", ts)
for a in sorted(agg, key=lambda a: -a[2])[:top]:
    print("%4d %6.2f%% inst %6.2f%% samp (long_sb %5.2f%%)  %s" % (a[0], 100 * a[3] / ti, 100 * a[2] / ts, 100 * a[4] / ts, a[1]))
