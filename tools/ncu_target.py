#!/usr/bin/env python
"""Small fixed workload for `ncu --set full` captures: one call of every heavy op at the
bench shapes (M=256) with a reduced row count so that the ~40 replays stay short."""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops  # noqa: E402

prec = ops.PREC[sys.argv[1] if len(sys.argv) > 1 else 'fp64']
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)


def rnd(*s):
    return torch.randn(*s, generator=g, dtype=torch.float64).to(dev)


M, D, Q, Do = 256, 10, 2, 2
x, z = rnd(n, D), rnd(M, D)
ls, sf = torch.full((D,), 0.5, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
A, B = rnd(Do, M), (0.01 * rnd(Do, M, M))
B = (B + B.transpose(1, 2)).contiguous()
dm, dv = rnd(n, Do), rnd(n, Do)
opnd = ops.DetOperands(prec, A, B)
for _ in range(2):
    m, v, Ks, Ts = ops.det_fwd(prec, x, z, ls, sf, opnd, save=True)
    ops.det_bwd(prec, x, z, ls, sf, opnd, dm, dv, Ks, Ts)
    ops.det_syrk(prec, Ks, dv, M)
mx, vx, z2 = rnd(n, Q), (0.1 + torch.rand(n, Q, dtype=torch.float64)).to(dev), rnd(M, Q)
ls2 = torch.full((Q,), 0.3, dtype=torch.float64, device=dev)
for _ in range(2):
    mo, vo, va, p1 = ops.mm_fwd(prec, mx, vx, z2, ls2, sf, A, B)
    ops.mm_bwd(prec, mx, vx, z2, ls2, sf, A, B, dm, dv, mo, va, p1)
from geepee_b200 import layers
Kuu = ops.kmat(z, z, ls, sf, 1e-5)
for _ in range(2):
    layers.spd_inverse(torch.stack([Kuu, Kuu + 0.1 * torch.eye(M, dtype=torch.float64, device=dev)]))
torch.cuda.synchronize()
print('ok')
