#!/usr/bin/env python
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops
from geepee_b200 import _lib
lib = _lib.get()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
z = torch.randn(M, 4, generator=g, dtype=torch.float64).to(dev)
ls = torch.zeros(4, dtype=torch.float64, device=dev); sf = torch.zeros(1, dtype=torch.float64, device=dev)
K = ops.kmat(z, z, ls, sf, 1e-5)
A = torch.stack([K + 0.1 * i * torch.eye(M, dtype=torch.float64, device=dev) for i in range(b)]).contiguous()
inv = torch.empty_like(A); ld = torch.zeros(128, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def f():
    rc = lib.gpb_spd_inverse(ctypes.c_void_p(A.data_ptr()), b, M, ctypes.c_void_p(inv.data_ptr()), ctypes.c_void_p(ld.data_ptr()), ctypes.c_void_p(st))
    assert rc == 0
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
err = (inv[0] @ A[0] - torch.eye(M, dtype=torch.float64, device=dev)).abs().max().item()
ref = torch.linalg.inv(A[0]); rel = ((inv[0] - ref).abs().max() / ref.abs().max()).item()
print('M=%d batch=%d spd_inverse %.1f us  |inv A - I|=%.2e rel vs torch %.2e logdet err %.2e' % (M, b, 1e3 * e0.elapsed_time(e1) / 20, err, rel, abs(ld[0].item() - torch.linalg.slogdet(A[0])[1].item())))
t = ld.cpu().numpy()
if t[16:].any():
    names = ['issue', 'stagewait', 'diag', 'sync', 'P', 'update', 'publish', 'csync']
    for r in range(4):
        print(' rank', r, 'tid0', ' '.join('%s %.0f' % (n, v) for n, v in zip(names, t[16 + 8 * r:24 + 8 * r])), '| tid200', ' '.join('%.0f' % v for v in t[48 + 8 * r:56 + 8 * r]))
