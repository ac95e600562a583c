#!/usr/bin/env python
"""ncu target: one forward + backward of a wide moment-matched layer (cfg2 shape, reduced rows)."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).to(dev)
n, M, Q, Do = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 128, 5, 50
mx, z = rnd(n, Q), rnd(M, Q)
vx = (0.1 + torch.rand(n, Q, dtype=torch.float64)).to(dev)
ls, sf = torch.full((Q,), 0.3, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
A, B = rnd(Do, M), (0.01 * rnd(Do, M, M)).contiguous()
dm, dv = rnd(n, Do), rnd(n, Do)
for _ in range(2):
    mo, vo, va, p1 = ops.mm_fwd(ops.F64, mx, vx, z, ls, sf, A, B)
    ops.mm_bwd(ops.F64, mx, vx, z, ls, sf, A, B, dm, dv, mo, va, p1)
torch.cuda.synchronize()
print('ok')
