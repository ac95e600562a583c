#!/usr/bin/env python
"""Digest of an .ncu-rep (all kernels): one block per kernel with the metrics the roofline
discussion in DESIGN.md uses, the warp-stall breakdown and the SASS instruction mix.
usage: ncu_digest.py report.ncu-rep > profiles/xxx.txt   (runs `ncu -i` to export the pages)"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]


def page(name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(l for l in out.splitlines() if not l.startswith('==')))


KEEP = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
raw = page('raw')
hdr, units = raw[0], raw[1]
ci = {h: i for i, h in enumerate(hdr)}
src = page('source')
heads = [i for i, r in enumerate(src) if len(r) > 3 and 'Source' in r and '# Samples' in r]
blocks = []
for hi, h in enumerate(heads):
    sh = src[h]
    cs = {x: i for i, x in enumerate(sh)}
    end = heads[hi + 1] - 1 if hi + 1 < len(heads) else len(src)
    body = [r for r in src[h + 1:end] if len(r) >= len(sh)]
    if not any('Instructions Executed' == x for x in sh):
        continue
    tot, ops = collections.Counter(), collections.Counter()
    ns = ni = 0
    for r in body:
        try:
            s = int(r[cs['# Samples']] or 0)
            ie = int(r[cs['Instructions Executed']] or 0)
        except ValueError:
            continue
        ns += s
        ni += ie
        w = r[cs['Source']].split()
        op = (w[1] if w and w[0].startswith('@') and len(w) > 1 else (w[0] if w else '')).split('.')[0]
        ops[op] += ie
        for x in sh:
            if x.startswith('stall_') and 'Not Issued' not in x:
                tot[x] += int(r[cs[x]] or 0)
    blocks.append((ns, ni, tot, ops))
# the source page lists every kernel twice (SASS / high-level view); keep one per kernel by
# matching the executed-instruction totals with the raw page
for k, r in enumerate(raw[2:]):
    print('== kernel %d: %s' % (k, r[ci['Kernel Name']]))
    for h in KEEP:
        if h in ci:
            print('  %-82s %s %s' % (h, r[ci[h]], units[ci[h]]))
    want = float(r[ci['smsp__inst_executed.sum']]) if 'smsp__inst_executed.sum' in ci else None
    for ns, ni, tot, ops in blocks:
        if want is not None and abs(ni - want) <= 1e-6 * max(want, 1):
            print('  warp stalls (%% of %d samples): %s' % (ns, ', '.join(
                '%s %.1f' % (a[6:], 100.0 * b / max(ns, 1)) for a, b in tot.most_common(8))))
            print('  instruction mix (%% of %d warp instructions): %s' % (ni, ', '.join(
                '%s %.1f' % (a, 100.0 * b / max(ni, 1)) for a, b in ops.most_common(12))))
            break
