#!/usr/bin/env python
"""Per-kernel stall / instruction-mix summary of an ncu `--page source --csv` export that holds
several kernels.  usage: ncu_src.py file.csv [kernel_index [n_top_lines]]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if len(r) > 3 and 'Source' in r and '# Samples' in r]
sel = int(sys.argv[2]) if len(sys.argv) > 2 else None
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for hi, h in enumerate(hdr):
    if sel is not None and hi != sel:
        continue
    sh = rows[h]
    ci = {x: i for i, x in enumerate(sh)}
    end = hdr[hi + 1] - 1 if hi + 1 < len(hdr) else len(rows)
    body = [r for r in rows[h + 1:end] if len(r) >= len(sh)]
    tot, ops = collections.Counter(), collections.Counter()
    ns = ni = 0
    for r in body:
        try:
            s = int(r[ci['# Samples']] or 0)
        except ValueError:
            continue
        ns += s
        ie = int(r[ci['Instructions Executed']] or 0)
        ni += ie
        w = r[ci['Source']].split()
        op = (w[1] if w and w[0].startswith('@') and len(w) > 1 else (w[0] if w else '')).split('.')[0]
        ops[op] += ie
        for x in sh:
            if x.startswith('stall_') and 'Not Issued' not in x:
                tot[x] += int(r[ci[x]] or 0)
    print('kernel', hi, 'samples', ns, 'warp-inst', ni)
    print('  stalls:', [(k, round(100 * v / max(ns, 1), 1)) for k, v in tot.most_common(9)])
    print('  mix:', [(k, round(100 * v / max(ni, 1), 1)) for k, v in ops.most_common(14)])
    if ntop:
        top = sorted(body, key=lambda r: -int(r[ci['# Samples']] or 0))[:ntop]
        for r in top:
            st = sorted(((int(r[ci[x]] or 0), x) for x in sh if x.startswith('stall_') and 'Not Issued' not in x), reverse=True)[:3]
            print('  %6s %-60s %s' % (r[ci['# Samples']], r[ci['Source']][:60], [(x[6:], v) for v, x in st]))
