#!/usr/bin/env python
"""Kernel timeline of one objective call (torch.profiler / CUPTI; no nsys in the image): where the GPU is idle or
runs only small kernels.  usage: step_trace.py [workload] [N] [prec]"""
import contextlib
import io
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geepee_b200 import aep_models as aep  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else 'cfg3_sdgpr'
w = dict(bench.WORKLOADS[wname])
if len(sys.argv) > 2:
    w['N'] = int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else 'fp64'
dev = torch.device('cuda', 0)
X, Y = bench.make_data(w)
with contextlib.redirect_stdout(io.StringIO()):
    if w['model'] == 'SGPR':
        model = aep.SGPR(X, Y, w['M'], prec=prec, device=dev)
    else:
        model = aep.SDGPR(X, Y, w['M'], w['hidden'], prec=prec, device=dev)
    params = bench.make_params(model, Y, w, X)
for _ in range(6):
    model.objective_function(params, w['N'], alpha=w['alpha'])
torch.cuda.synchronize()
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
host = 0.0
for _ in range(10):
    h0 = time.perf_counter()
    model.objective_function(params, w['N'], alpha=w['alpha'])
    host += time.perf_counter() - h0
e1.record()
torch.cuda.synchronize()
print('untraced: %.3f ms per step by CUDA events, %.3f ms per call on the host clock' % (e0.elapsed_time(e1) / 10, 1e3 * host / 10))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        model.objective_function(params, w['N'], alpha=w['alpha'])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# the last step: after the last long gap between memcpy HtoD events ... simpler: take the last third of the kernels
t_first = ev[0].time_range.start
n3 = len(ev) // 3
step = ev[2 * n3:]
t0 = step[0].time_range.start
t1 = max(e.time_range.end for e in step)
print('%s N=%d %s: %d GPU activities per step, step span %.3f ms' % (wname, w['N'], prec, len(step), (t1 - t0) / 1e3))
BIG = 200.0      # us
busy_big = sum(e.time_range.end - e.time_range.start for e in step if e.time_range.end - e.time_range.start >= BIG)
print('kernels >= %d us: %.3f ms; everything else (exposed or hidden): span - big = %.3f ms' % (BIG, busy_big / 1e3, (t1 - t0 - busy_big) / 1e3))
# timeline: merge the big kernels into segments, describe what happens in the gaps between them
bigs = [e for e in step if e.time_range.end - e.time_range.start >= BIG]
cur = t0
for b in bigs + [None]:
    end = b.time_range.start if b is not None else t1
    gap = end - cur
    if gap > 15.0:
        inside = [e for e in step if e.time_range.start >= cur - 1 and e.time_range.end <= end + 1 and e not in bigs]
        names = {}
        for e in inside:
            k = e.name[:48]
            d = names.setdefault(k, [0, 0.0])
            d[0] += 1
            d[1] += e.time_range.end - e.time_range.start
        top = sorted(names.items(), key=lambda kv: -kv[1][1])[:6]
        print('  gap %8.1f us at +%9.1f us before %-40s: %d small activities, %s' % (
            gap, cur - t0, (b.name[:40] if b is not None else 'END'), len(inside),
            '; '.join('%s x%d %.0fus' % (k, c, t) for k, (c, t) in top)))
    if b is not None:
        cur = max(cur, b.time_range.end)
if os.environ.get('GPB_TRACE_HEAD'):
    nh = int(os.environ['GPB_TRACE_HEAD'])
    print('first %d activities of the step (start offset us, duration us, name):' % nh)
    for e in step[:nh]:
        print('  %9.1f %8.1f  %s' % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
    print('last %d activities:' % nh)
    for e in step[-nh - 25:]:
        print('  %9.1f %8.1f  %s' % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name[:90]))
