#!/usr/bin/env python
"""Share of each kernel in the LAST step of an `ncu --metrics gpu__time_duration.sum --csv` launch
list of `bench.py --steps 1 --warmup W` (the steps are delimited by the packed parameter upload's
first library kernel; simpler and robust: split the launch list into W+1(+e2e) equal-structure
groups by the det_fwd launches and summarise the group before the e2e pass).
usage: launch_share.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
recs = [(r[ci['Kernel Name']], float(r[ci['Metric Value']].replace(',', '')), r[ci['Metric Unit']])
        for r in rows[1:] if len(r) > ci['Metric Value'] and r[ci['Metric Name']] == 'gpu__time_duration.sum']
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}
recs = [(k, v * scale.get(u, 1e-6)) for k, v, u in recs]
starts = [i for i, (k, _) in enumerate(recs) if 'det_fwd' in k]
print('launches %d, steps found %d' % (len(recs), len(starts)))
if len(starts) >= 2:
    # the objective's first library launches (kmat / pre-tail) precede det_fwd; take whole periods
    period = starts[-1] - starts[-2]
    lo = starts[-2]
    step = recs[lo:lo + period]
    tot = sum(v for _, v in step)
    agg = collections.Counter()
    for k, v in step:
        name = k.split('(')[0].split('<')[0].replace('void ', '').replace('gpb::', '')
        agg[name] += v
    print('one step: %d launches, %.3f ms of kernel time (serialised, cold cache)' % (len(step), tot))
    lib = sum(v for k, v in agg.items() if not k.startswith(('at::', 'cublas', 'cutlass', 'sm', 'void at')) and 'at::native' not in k)
    for k, v in agg.most_common(14):
        print('  %-48s %9.3f ms  %5.1f %%' % (k[:48], v, 100 * v / tot))
