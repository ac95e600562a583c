"""`torch.ops.geepee_b200.*`: the tensor-in / tensor-out entry points of the library registered as PyTorch custom ops
(BASELINE.json north_star: "Python host code calls hand-written sm_100a CUDA kernels through PyTorch custom ops").

    import geepee_b200.torch_ops                       # registers the namespace
    psi1, psi2 = torch.ops.geepee_b200.psi_stats(mx, vx, z, ls, sf)

Each op is the ctypes call of `geepee_b200.ops` (one C-ABI call into libgeepee_b200.so on torch's current stream)
under a dispatcher schema with a shape function (`register_fake`), so the ops can be traced, exported and called from
C++ / TorchScript hosts by name.  The model classes themselves keep calling `geepee_b200.ops` directly: the dispatcher
adds microseconds per call, which the launch-bound configs (cfg1: 68 launches in 0.6 ms) would feel, and buys nothing
there.  There is no CPU implementation behind these names: on a box without the CUDA library the call fails in
`_lib.get()` like every other entry point.

reference lines: kmat kernels.py:10-22; psi_stats kernels.py:181-240; spd_inverse base_models.py:464-476;
gauss_lik lik_layers.py:104-133,183-199; mm_fwd aep_models.py:183-199; mm_bwd aep_models.py:238-250 +
kernels.py:302-309,355-378,402-444."""
from typing import Tuple

import torch
from torch import Tensor

from . import ops

NS = 'geepee_b200'
_f64 = torch.float64


def _e(like, *shape):
    return torch.empty(shape, dtype=_f64, device=like.device)


@torch.library.custom_op(NS + '::kmat', mutates_args=())
def kmat(x: Tensor, z: Tensor, ls: Tensor, sf: Tensor, jitter: float = 0.0) -> Tensor:
    return ops.kmat(x, z, ls, sf, jitter)


@kmat.register_fake
def _(x, z, ls, sf, jitter=0.0):
    return _e(x, x.shape[0], z.shape[0])


@torch.library.custom_op(NS + '::psi_stats', mutates_args=())
def psi_stats(mx: Tensor, vx: Tensor, z: Tensor, ls: Tensor, sf: Tensor) -> Tuple[Tensor, Tensor]:
    return ops.psi_stats(mx, vx, z, ls, sf)


@psi_stats.register_fake
def _(mx, vx, z, ls, sf):
    n, M = mx.shape[0], z.shape[0]
    return _e(mx, n, M), _e(mx, n, M, M)


@torch.library.custom_op(NS + '::spd_inverse', mutates_args=())
def spd_inverse(A: Tensor) -> Tuple[Tensor, Tensor]:
    inv, ld = ops.spd_inverse(A)
    return inv.clone() if A.dim() == 2 else inv, ld.clone() if A.dim() == 2 else ld


@spd_inverse.register_fake
def _(A):
    if A.dim() == 2:
        return torch.empty_like(A), _e(A)
    return torch.empty_like(A), _e(A, A.shape[0])


@torch.library.custom_op(NS + '::gauss_lik', mutates_args=())
def gauss_lik(m: Tensor, v: Tensor, y: Tensor, sn: Tensor, alpha: float, scale: float,
              mode: int) -> Tuple[Tensor, Tensor, Tensor]:
    return ops.gauss_lik(m, v, y, sn, alpha, scale, mode)


@gauss_lik.register_fake
def _(m, v, y, sn, alpha, scale, mode):
    return torch.empty_like(m), torch.empty_like(m), _e(m, 2)


@torch.library.custom_op(NS + '::mm_fwd', mutates_args=())
def mm_fwd(prec: int, mx: Tensor, vx: Tensor, z: Tensor, ls: Tensor, sf: Tensor, A: Tensor,
           B: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> mout, vout, vacc, psi1 (the last two are what mm_bwd reuses)."""
    return ops.mm_fwd(prec, mx, vx, z, ls, sf, A, B, save=True)


@mm_fwd.register_fake
def _(prec, mx, vx, z, ls, sf, A, B):
    n, Do, M = mx.shape[0], A.shape[0], A.shape[1]
    return _e(mx, n, Do), _e(mx, n, Do), _e(mx, n, Do), _e(mx, n, M)


MM_BWD_OUTPUTS = ('dA', 'dB', 'dzu', 'dl', 'dsf2', 'dvsum', 'dmx', 'dvx')


@torch.library.custom_op(NS + '::mm_bwd', mutates_args=())
def mm_bwd(prec: int, mx: Tensor, vx: Tensor, z: Tensor, ls: Tensor, sf: Tensor, A: Tensor, B: Tensor, dm: Tensor,
           dv: Tensor, mout: Tensor, vacc: Tensor,
           psi1: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> the tensors named in MM_BWD_OUTPUTS, in that order."""
    out = ops.mm_bwd(prec, mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1)
    return tuple(out[k] for k in MM_BWD_OUTPUTS)


@mm_bwd.register_fake
def _(prec, mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1):
    n, Q, Do, M = mx.shape[0], mx.shape[1], A.shape[0], A.shape[1]
    return (_e(mx, Do, M), _e(mx, Do, M, M), _e(mx, M, Q), _e(mx, Q), _e(mx, 1), _e(mx, 1), _e(mx, n, Q),
            _e(mx, n, Q))
