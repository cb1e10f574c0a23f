#!/usr/bin/env python3
"""usage: hot_lines.py REPORT.ncu-rep [N] [LAUNCH]  -> the N CUDA source lines with the most warp-stall samples of launch LAUNCH
(default 0) of the report (needs -lineinfo and --import-source on): share of the stall samples, of the executed warp instructions,
of the L1 tag requests, and the two dominant memory stall reasons (long scoreboard, LG throttle) as shares of all samples."""
import csv
import io
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out[out.find('"'):])))
cur, lines, col = None, [], {}


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        col = {h: i for i, h in enumerate(r)}
    elif len(r) > 7 and r[0].isdigit():
        g = lambda name: f(r[col[name]]) if name in col and col[name] < len(r) else 0.0
        lines.append((cur, int(r[0]), f(r[4]), f(r[7]), r[1].strip(), g("L1 Tag Requests Global"), g("stall_long_sb"), g("stall_lg")))
ts, ti, tr = sum(x[2] for x in lines) or 1, sum(x[3] for x in lines) or 1, sum(x[5] for x in lines) or 1
print("file:line | stall samples | executed warp instr | L1 tag requests | long_sb | lg_throttle | source")
for c, ln, s, i, src, rq, lsb, lg in sorted(lines, key=lambda x: -x[2])[:top]:
    print("%s:%d | %.3f | %.3f | %.3f | %.3f | %.3f | %s" % (c, ln, s / ts, i / ti, rq / tr, lsb / ts, lg / ts, src[:110]))
