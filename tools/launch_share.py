#!/usr/bin/env python
"""Per-kernel share of ONE objective call from an `ncu --metrics gpu__time_duration.sum --csv` launch
list of `bench.py --steps 1 --warmup W --no-cpu` (every objective call launches the same kernel
sequence -- eagerly or from CUDA-graph replays -- so the step is the shortest period of the name
sequence once the trailing FMA-peak probe is dropped).
usage: launch_share.py launches.csv [--all]"""
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
while recs and ('fma_peak' in recs[-1][0] or 'FillFunctor' in recs[-1][0]):
    recs.pop()
names = [k for k, _ in recs]
period = None
for p in range(8, len(names) // 2 + 1):
    if names[-p:] == names[-2 * p:-p]:
        period = p
        break
print('launches %d, period %s' % (len(recs), period))
if period:
    step = recs[-period:]
    tot = sum(v for _, v in step)

    def short(k):
        return k.split('(')[0].replace('void ', '').replace('gpb::', '')[:70]
    agg, cnt = collections.Counter(), collections.Counter()
    for k, v in step:
        agg[short(k)] += v
        cnt[short(k)] += 1
    foreign = [k for k in agg if k.startswith(('at::', 'cublas', 'cutlass', 'sm80', 'sm90', 'sm100', 'nccl')) or 'at::native' in k
               or 'cublas' in k.lower() or 'gemv' in k.lower() or 'elementwise' in k]
    print('one step: %d launches, %.3f ms of kernel time (serialised, cold cache); not from this library: %d launches %.3f ms'
          % (len(step), tot, sum(cnt[k] for k in foreign), sum(agg[k] for k in foreign)))
    lim = 1000 if '--all' in sys.argv else 24
    for k, v in agg.most_common(lim):
        print('  %-70s x%-3d %9.4f ms  %5.1f %%%s' % (k, cnt[k], v, 100 * v / tot, '   <-- foreign' if k in foreign else ''))
