#!/usr/bin/env python
"""Own cluster Gauss-Jordan SPD inverse vs the torch (cuSOLVER potrf + trsm + GEMM) path."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops

dev = torch.device('cuda:0')


def torch_path(A):
    L, _ = torch.linalg.cholesky_ex(A, check_errors=False)
    eye = torch.eye(A.shape[-1], dtype=A.dtype, device=A.device).expand(A.shape[0], -1, -1)
    Linv = torch.linalg.solve_triangular(L, eye, upper=False)
    return torch.matmul(Linv.transpose(-1, -2), Linv), 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for M, b in [(128, 5), (200, 9), (256, 1), (256, 5), (512, 3)]:
    g = torch.Generator().manual_seed(M)
    z = torch.randn(b, M, 4, generator=g, dtype=torch.float64)
    d2 = ((z[:, :, None, :] - z[:, None, :, :])**2).sum(-1)
    A = (torch.exp(-0.5 * d2) + 1e-5 * torch.eye(M, dtype=torch.float64)).to(dev).contiguous()
    i1, l1 = ops.spd_inverse(A)
    i2, l2 = torch_path(A)
    err = float((i1 - i2).abs().max() / i2.abs().max())
    res = float((torch.matmul(i1, A) - torch.eye(M, dtype=torch.float64, device=dev)).abs().max())
    res2 = float((torch.matmul(i2, A) - torch.eye(M, dtype=torch.float64, device=dev)).abs().max())
    print('M=%d batch=%d: own %.3f ms, torch %.3f ms; rel diff %.2e, residual own %.2e torch %.2e, logdet diff %.2e'
          % (M, b, timeit(lambda: ops.spd_inverse(A)), timeit(lambda: torch_path(A)), err, res, res2,
             float((l1 - l2).abs().max())))
