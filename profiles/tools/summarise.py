#!/usr/bin/env python3
"""usage: summarise.py REPORT.ncu-rep [REPORT...]  -> markdown on stdout.
Reads `ncu -i REPORT --page raw --csv` (selected metrics per launch) and `--page source --csv` (opcode mix and
stall-sample shares per opcode), as recommended in /opt/skills/guides/B200_PROFILING.md."""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg",
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle", "not_selected",
          "selected", "branch_resolving", "barrier", "dispatch_stall", "no_instruction"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    start = out.find('"')
    return list(csv.reader(io.StringIO(out[start:])))


def raw(rep):
    rows = ncu(rep, "raw")
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    name_col = idx.get("Kernel Name")
    launches = rows[2:]
    print("| metric | unit | " + " | ".join("launch %d (%s)" % (i, (r[name_col].split("(")[0] if name_col is not None else "")[:40]) for i, r in enumerate(launches)) + " |")
    print("|---|---|" + "---|" * len(launches))
    for m in METRICS:
        if m in idx:
            print("| %s | %s | " % (m, units[idx[m]]) + " | ".join(r[idx[m]] for r in launches) + " |")
    # stall reasons: warp-sample shares
    for r in launches:
        tot, vals = 0.0, {}
        for s in STALLS:
            k = "smsp__pcsamp_warps_issue_stalled_%s" % s
            if k in idx:
                try:
                    vals[s] = float(r[idx[k]].replace(",", ""))
                    tot += vals[s]
                except ValueError:
                    pass
        if tot:
            print("\nstall samples (share): " + ", ".join("%s %.3f" % (s, v / tot) for s, v in vals.items()))


def source(rep):
    rows = ncu(rep, "source")
    while rows and "Source" not in rows[0]:
        rows = rows[1:]                      # one "Kernel Name" line precedes the header
    if not rows:
        return
    head = rows[0]
    idx = {h: i for i, h in enumerate(head)}
    src = idx.get("Source")
    ex = idx.get("Instructions Executed")
    sm = idx.get("# Samples", idx.get("Warp Stall Sampling (All Samples)"))
    if src is None or ex is None:
        return
    n, e, s = 0, collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r and r[0] == "Kernel Name":
            break                            # next launch of a multi-launch report
        if len(r) <= max(src, ex):
            continue
        toks = [t for t in r[src].split() if not t.startswith("@")]
        if not toks:
            continue
        op = toks[0].split(".")[0]
        try:
            e[op] += float(r[ex].replace(",", "") or 0)
            if sm is not None:
                s[op] += float(r[sm].replace(",", "") or 0)
        except ValueError:
            continue
        n += 1
    te, ts = sum(e.values()) or 1, sum(s.values()) or 1
    print("\nSASS lines: %d; opcode mix (share of executed warp instructions / share of stall samples):\n" % n)
    print(", ".join("%s %.3f/%.3f" % (op, v / te, s[op] / ts) for op, v in e.most_common(14)))


for rep in sys.argv[1:]:
    print("\n## %s\n" % rep.split("/")[-1])
    raw(rep)
    source(rep)
