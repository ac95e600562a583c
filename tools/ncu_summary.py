#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages exported as CSV) into a small text table."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
print('metric,unit,value')
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print('%s,%s,%s' % (h, u, v))
srows = list(csv.reader(open(src)))
sh = srows[1]
ci = {h: i for i, h in enumerate(sh)}
tot, ops = collections.Counter(), collections.Counter()
ns = ni = 0
for r in srows[2:]:
    if len(r) < len(sh):
        continue
    s = int(r[ci['# Samples']] or 0)
    ns += s
    ie = int(r[ci['Instructions Executed']] or 0)
    ni += ie
    w = r[ci['Source']].split()
    op = (w[1] if w and w[0].startswith('@') and len(w) > 1 else (w[0] if w else '')).split('.')[0]
    ops[op] += ie
    for h in sh:
        if h.startswith('stall_') and 'Not Issued' not in h:
            tot[h] += int(r[ci[h]] or 0)
print('# warp stall sampling (all samples = %d), SASS instructions = %d' % (ns, len(srows) - 2))
for k, v in tot.most_common(8):
    print('%s,pct,%.1f' % (k, 100.0 * v / max(ns, 1)))
print('# instruction mix (warp instructions executed = %d)' % ni)
for k, v in ops.most_common(12):
    print('inst_%s,pct,%.1f' % (k, 100.0 * v / max(ni, 1)))
