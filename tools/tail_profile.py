#!/usr/bin/env python
"""Where does the non-kernel time of one objective call go?  torch.profiler over a few steps of
the SGPR north-star shape (kernels ~14 ms, everything else is 'tail' + launch overhead)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench
from geepee_b200 import aep_models as aep

wl = sys.argv[1] if len(sys.argv) > 1 else 'ns_sgpr'
w = dict(bench.WORKLOADS[wl])
if len(sys.argv) > 2:
    w['N'] = int(sys.argv[2])
X, Y = bench.make_data(w)
dev = torch.device('cuda:0')
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    if w['model'] == 'SGPR':
        model = aep.SGPR(X, Y, w['M'], device=dev)
    elif w['model'] == 'SGPLVM':
        model = aep.SGPLVM(Y, w['Q'], w['M'], device=dev)
    elif w['model'] == 'SGPSSM':
        model = aep.SGPSSM(Y, w['Q'], w['M'], device=dev)
    else:
        model = aep.SDGPR(X, Y, w['M'], w['hidden'], device=dev)
    params = bench.make_params(model, Y, w, X)
for _ in range(3):
    model.objective_function(params, w['N'], alpha=w['alpha'])
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    model.objective_function(params, w['N'], alpha=w['alpha'])
torch.cuda.synchronize()
print('wall ms/step', (time.perf_counter() - t) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        model.objective_function(params, w['N'], alpha=w['alpha'])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by='cpu_time_total', row_limit=20, max_name_column_width=60))
