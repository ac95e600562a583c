#!/usr/bin/env python
"""cProfile of the host side of objective_function (where the launch-bound configs spend their step)."""
import contextlib
import cProfile
import io
import os
import pstats
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from geepee_b200 import aep_models as aep  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else 'cfg1_sgpr'
w = dict(bench.WORKLOADS[wname])
if len(sys.argv) > 2:
    w['N'] = int(sys.argv[2])
dev = torch.device('cuda', 0)
X, Y = bench.make_data(w)
with contextlib.redirect_stdout(io.StringIO()):
    if w['model'] == 'SGPR':
        model = aep.SGPR(X, Y, w['M'], device=dev)
    else:
        model = aep.SDGPR(X, Y, w['M'], w['hidden'], device=dev)
    params = bench.make_params(model, Y, w, X)
for _ in range(10):
    model.objective_function(params, w['N'], alpha=w['alpha'])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    model.objective_function(params, w['N'], alpha=w['alpha'])
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats('cumulative').print_stats(45)
print(st.getvalue()[:9000])
