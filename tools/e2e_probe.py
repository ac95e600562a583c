#!/usr/bin/env python
"""Does the double-buffered upload overlap the step?  NS shape, per-step times of: resident, serial copy, prefetch."""
import contextlib, io, os, sys, time
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench
from geepee_b200 import aep_models as aep
w = bench.WORKLOADS['ns_sgpr']
dev = torch.device('cuda', 0)
X, Y = bench.make_data(w)
with contextlib.redirect_stdout(io.StringIO()):
    model = aep.SGPR(X, Y, w['M'], device=dev)
    params = bench.make_params(model, Y, w, X)
N = w['N']
xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
def step():
    return model.objective_function(params, N, alpha=w['alpha'])
def timeit(fn, k=10):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print('resident      %.3f ms' % timeit(step))
def copy_only():
    model._x.copy_(xh, non_blocking=True); model._y.copy_(yh, non_blocking=True)
print('copy only     %.3f ms' % timeit(copy_only))
def serial():
    copy_only(); return step()
print('serial copy   %.3f ms' % timeit(serial))
cs = torch.cuda.Stream(dev)
x2, y2 = torch.empty_like(model._x), torch.empty_like(model._y)
def overlapped():
    with torch.cuda.stream(cs):
        x2.copy_(xh, non_blocking=True); y2.copy_(yh, non_blocking=True)
    return step()
print('copy on a side stream into an unrelated buffer + step   %.3f ms' % timeit(overlapped))
