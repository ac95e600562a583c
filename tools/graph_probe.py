#!/usr/bin/env python
"""Can the layer pre-tail (kmat + Cholesky-based inverses + GEMMs, geepee_b200/layers.py) be
captured in a CUDA graph on a side stream and replayed?  Development probe."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import bench
from geepee_b200 import aep_models as aep, layers

dev = torch.device('cuda:0')
w = dict(bench.WORKLOADS['cfg3_sdgpr']); w['N'] = 4096
X, Y = bench.make_data(w)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    model = aep.SDGPR(X, Y, w['M'], w['hidden'], device=dev)
    params = bench.make_params(model, Y)
L = model.sgp_layers[1]
pdev = layers.pack_to_device(params, dev)
static = {k: v.clone() for k, v in pdev.items()}


def pre():
    L._fuse_cavity_alpha = 1.0
    L.update_hypers(params, key_suffix='_1', _dev=static)
    L.compute_cavity(1.0)


pre(); torch.cuda.synchronize()
ref = {k: v.clone() for k, v in L._t.items()}
t = time.perf_counter()
for _ in range(20): pre()
torch.cuda.synchronize()
print('eager pre-tail ms', (time.perf_counter() - t) / 20 * 1e3)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.stream(s):
        for _ in range(3): pre()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        pre()
    torch.cuda.synchronize()
    outs = dict(L._t)
    for k in static: static[k].mul_(1.0)
    g.replay(); torch.cuda.synchronize()
    err = max(float((outs[k] - ref[k]).abs().max()) for k in ref if k in outs)
    print('graph replay max abs diff vs eager', err)
    t = time.perf_counter()
    for _ in range(20): g.replay()
    torch.cuda.synchronize()
    print('graph pre-tail ms', (time.perf_counter() - t) / 20 * 1e3)
except Exception as e:
    print('capture failed:', type(e).__name__, str(e)[:400])
