#!/usr/bin/env python3
"""usage: summarise.py [--json OUT.json] REPORT.ncu-rep [REPORT...]  -> markdown on stdout.
--json additionally writes the counters bench.py quotes (DRAM bytes, pipe utilisation, L1/L2 figures, duration) per
kernel of the first launch of every report, together with the git revision, to OUT.json (tracked under profiles/rNN/).
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


JSON_OUT = {}


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def raw(rep):
    rows = ncu(rep, "raw")
    head, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(head)}
    name_col = idx.get("Kernel Name")
    launches = rows[2:]
    for r in launches:
        name = r[name_col] if name_col is not None else rep
        if name in JSON_OUT and JSON_OUT[name]["report"] != rep.split("/")[-1]:
            continue
        if name in JSON_OUT:                # further launches of the same kernel in one report (e.g. search pass 1, pass 2)
            k = 1
            while "%s #launch %d" % (name, k) in JSON_OUT:
                k += 1
            name = "%s #launch %d" % (name, k)

        def num(m, scale=1.0):
            try:
                return round(float(r[idx[m]].replace(",", "")) * scale, 3)
            except (KeyError, ValueError):
                return None
        try:
            dram = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                   to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        except (KeyError, ValueError):
            dram = None
        JSON_OUT[name] = {
            "report": rep.split("/")[-1], "dram_bytes": dram,
            "duration_us_under_ncu": num("gpu__time_duration.sum"),
            "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1tex_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num("lts__t_sector_hit_rate.pct"),
            "l1_sectors_global_ld": num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "warp_instructions": num("smsp__inst_executed.sum"), "registers": num("launch__registers_per_thread")}
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


args = sys.argv[1:]
json_path = None
if args and args[0] == "--json":
    json_path, args = args[1], args[2:]
for rep in args:
    print("\n## %s\n" % rep.split("/")[-1])
    raw(rep)
    source(rep)
if json_path:
    import json
    import os
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    prev = {}
    if os.path.exists(json_path):
        prev = json.load(open(json_path)).get("kernels", {})
    prev.update(JSON_OUT)
    json.dump({"git": git, "how": "ncu --set full --clock-control none, first captured launch of each report; profiles/tools/summarise.py --json",
               "kernels": prev}, open(json_path, "w"), indent=1)
