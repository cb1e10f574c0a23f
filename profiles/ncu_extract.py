"""Helper used to turn .ncu-rep files (gpurun_out/, scratch) into the committed summaries."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg']
STALLS = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_lg', 'stall_not_selected', 'stall_selected',
          'stall_branch_resolving', 'stall_barrier', 'stall_dispatch', 'stall_no_inst']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []
            secs.append(cur)
        elif cur is not None:
            cur.append(r)
    return secs


def summarize(rep, title, f):
    hdr, units, data = raw(rep)
    f.write(f'## {title}\n\n| metric | unit | ' + ' | '.join(f'launch {i}' for i in range(len(data))) + ' |\n|---|---|' + '---|' * len(data) + '\n')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            f.write(f'| {k} | {units[i]} | ' + ' | '.join(r[i] for r in data) + ' |\n')
    sec = source(rep)[-1]
    h, d = sec[0], sec[1:]
    ii, si, so = h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
    ops, samp = collections.Counter(), collections.Counter()
    for r in d:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[so])
        op = m.group(2).split('.')[0] if m else '?'
        ops[op] += int(r[ii]); samp[op] += int(r[si])
    tot, ts = sum(ops.values()), max(sum(samp.values()), 1)
    f.write(f'\nSASS lines: {len(d)}; opcode mix (share of executed warp instructions / share of stall samples):\n\n')
    f.write(', '.join(f'{op} {c / tot:.3f}/{samp[op] / ts:.3f}' for op, c in ops.most_common(14)) + '\n\n')
    f.write('stall reasons (share of samples): ' + ', '.join(f'{n} {sum(int(r[h.index(n)]) for r in d) / ts:.3f}' for n in STALLS if n in h) + '\n\n')


if __name__ == '__main__':
    out = sys.argv[1]
    with open(out, 'a') as f:
        for spec in sys.argv[2:]:
            rep, title = spec.split('=', 1)
            summarize(rep, title, f)
