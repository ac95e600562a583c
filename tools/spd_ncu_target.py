import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops
dev = torch.device('cuda:0')
M, b = 256, 5
g = torch.Generator().manual_seed(M)
z = torch.randn(b, M, 4, generator=g, dtype=torch.float64)
d2 = ((z[:, :, None, :] - z[:, None, :, :])**2).sum(-1)
A = (torch.exp(-0.5 * d2) + 1e-5 * torch.eye(M, dtype=torch.float64)).to(dev).contiguous()
for _ in range(3):
    ops.spd_inverse(A)
torch.cuda.synchronize()
