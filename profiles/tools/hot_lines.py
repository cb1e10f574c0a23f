#!/usr/bin/env python3
"""usage: hot_lines.py REPORT.ncu-rep [N]  -> the N CUDA source lines with the most warp-stall samples (needs -lineinfo
and --import-source on), with their share of executed warp instructions."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.find('"'):])))
cur, lines = None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 7 and r[0].isdigit():
        def f(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.0
        lines.append((cur, int(r[0]), f(r[4]), f(r[7]), r[1].strip()))
ts, ti = sum(x[2] for x in lines) or 1, sum(x[3] for x in lines) or 1
print("file:line | stall samples | executed warp instr | source")
for c, ln, s, i, src in sorted(lines, key=lambda x: -x[2])[:top]:
    print("%s:%d | %.3f | %.3f | %s" % (c, ln, s / ts, i / ti, src[:120]))
