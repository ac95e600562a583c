#!/usr/bin/env python
"""Time the fp32 moment-matched forward alone (CUDA events) at the bench shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
Do = int(sys.argv[2]) if len(sys.argv) > 2 else 2
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 2
M = int(sys.argv[4]) if len(sys.argv) > 4 else 256
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g, dtype=torch.float64).to(dev)


mx, vx, z = rnd(n, Q), (0.1 + torch.rand(n, Q, generator=g, dtype=torch.float64)).to(dev), rnd(M, Q)
ls = torch.full((Q,), 0.3, dtype=torch.float64, device=dev)
sf = torch.zeros(1, dtype=torch.float64, device=dev)
A, B = rnd(Do, M), 0.01 * rnd(Do, M, M)
B = (B + B.transpose(1, 2)).contiguous()


def timeit(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name in ('fp32', 'fp64'):
    prec = ops.PREC[name]
    t = timeit(lambda: ops.mm_fwd(prec, mx, vx, z, ls, sf, A, B))
    out = ops.mm_fwd(prec, mx, vx, z, ls, sf, A, B)
    if name == 'fp32':
        o32 = out
    else:
        err = max(((o32[i] - out[i]).abs().max() / out[i].abs().max()).item() for i in range(3))
        print('fp32 vs fp64 max rel err (mout, vout, vacc): %.2e' % err)
    pairs = M * (M + 1) / 2
    print('%s n=%d Q=%d Do=%d M=%d mm_fwd %.3f ms  (%.2f G exp/s)' % (name, n, Q, Do, M, t, n * pairs / t / 1e6))
