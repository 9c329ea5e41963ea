#!/usr/bin/env python
"""Summarise an ncu report of k_substeps: headline metrics, stall mix, and instructions / stall samples per
source function (needs -lineinfo + --import-source on).  usage: python tools/ncu_summary.py REPORT.ncu-rep [out.md]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]


def ncu(*args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu('--page', 'raw', '--csv'))))
hdr, units = raw[0], raw[1]
out = []
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum']
for row in raw[2:]:
    out.append('## launch: %s' % row[hdr.index('Kernel Name')])
    for k in KEYS:
        if k in hdr:
            out.append('%-62s %14s %s' % (k, row[hdr.index(k)], units[hdr.index(k)]))

src = list(csv.reader(io.StringIO(ncu('--page', 'source', '--csv', '--print-source', 'sass,cuda'))))
fname, cols = None, None
per_line = defaultdict(lambda: [0, 0, defaultdict(int)])     # (file, line) -> [inst, samples, stalls]
stall_tot = defaultdict(int)
for r in src:
    if len(r) == 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        cols = r
        continue
    if cols is None or len(r) < len(cols) or not r[0]:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        inst = int(r[cols.index('Instructions Executed')] or 0)
        smp = int(r[cols.index('# Samples')] or 0)
    except ValueError:      # a source line whose quoting broke the CSV row
        continue
    e = per_line[(fname, line)]
    e[0] += inst
    e[1] += smp
    for i, c in enumerate(cols):
        if c.startswith('stall_') and 'Not Issued' not in c and r[i].isdigit():
            e[2][c] += int(r[i]); stall_tot[c] += int(r[i])

# map lines of b2s_step.cu to enclosing function
funcs = []
try:
    text = open('robovat_b200/csrc/b2s_step.cu').read().split('\n')
    pat = re.compile(r'^(?:__device__|__global__)[^;]*?\b(\w+)\s*\(')
    for i, l in enumerate(text, 1):
        m = pat.match(l)
        if m and not l.rstrip().endswith(';'):
            funcs.append((i, m.group(1)))
except IOError:
    pass


def func_of(f, line):
    if f != 'b2s_step.cu':
        return f
    name = '?'
    for s, n in funcs:
        if s <= line:
            name = n
    return name


agg = defaultdict(lambda: [0, 0])
for (f, line), (inst, smp, st) in per_line.items():
    a = agg[func_of(f, line)]
    a[0] += inst; a[1] += smp
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
out.append('\n## warp instructions and stall samples by source function (inlined callees count where their lines live)')
out.append('%-28s %14s %7s %10s %7s' % ('function/file', 'warp insts', '%', 'samples', '%'))
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append('%-28s %14d %6.1f%% %10d %6.1f%%' % (k, i, 100.0 * i / ti, s, 100.0 * s / ts))
out.append('\n## stall reasons over all samples')
tt = sum(stall_tot.values()) or 1
for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]):
    if v:
        out.append('%-28s %10d %6.1f%%' % (k, v, 100.0 * v / tt))
out.append('\n## hottest source lines (by samples)')
for (f, line), (inst, smp, st) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:40]:
    top = max(st.items(), key=lambda kv: kv[1])[0] if st else '-'
    out.append('%-14s:%-5d %-24s inst %12d samples %8d  top %s' % (f, line, func_of(f, line), inst, smp, top))
text = '\n'.join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text + '\n')
