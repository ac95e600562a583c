#!/usr/bin/env python
"""Time the fp32 deterministic-layer kernels alone (CUDA events) at the bench shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
Do = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = ops.PREC['fp32']
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g, dtype=torch.float64).to(dev)


M, D = 256, 10
x, z = rnd(n, D), rnd(M, D)
ls, sf = torch.full((D,), 0.5, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
A, B = rnd(Do, M), (0.01 * rnd(Do, M, M))
B = (B + B.transpose(1, 2)).contiguous()
dv = rnd(n, Do)
opnd = ops.DetOperands(prec, A, B)


def timeit(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t = timeit(lambda: ops.det_fwd(prec, x, z, ls, sf, opnd, save=True))
m, v, Ks, Ts = ops.det_fwd(prec, x, z, ls, sf, opnd, save=True)
t2 = timeit(lambda: ops.det_syrk(prec, Ks, dv, M))
print('dbg=%s n=%d Do=%d det_fwd %.3f ms (%.2f us/tile-slot)  det_syrk %.3f ms' % (
    os.environ.get('GPB_UMMA_DBG', '0'), n, Do, t, 1e3 * t / ((n / 128 + 147) // 148), t2))
