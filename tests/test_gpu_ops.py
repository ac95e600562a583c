"""Kernel-level parity on the B200: every C-ABI op of libgeepee_b200.so against the oracle's
numpy formulas (same bodies as tests/test_emu_ops.py), at small shapes plus shapes that span
several tiles / blocks / row splits."""
import pytest
import torch

import ops_cases as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()          # raises if libgeepee_b200.so is missing: no fallback
    oc.DEV = 'cuda'
    yield
    oc.DEV = 'cpu'


GPU_DET = oc.DET_SHAPES + [(5000, 200, 10, 2), (3000, 512, 16, 1), (20000, 50, 1, 1)]
GPU_MM = oc.MM_SHAPES + [(700, 64, 5, 3), (2000, 40, 2, 2), (300, 128, 4, 5)]


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-11), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,D,Do', GPU_DET)
def test_det_layer(n, M, D, Do, prec, tol):
    oc.check_det_layer(n, M, D, Do, prec, tol)


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-11), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,Q,Do', GPU_MM)
def test_mm_layer(n, M, Q, Do, prec, tol):
    oc.check_mm_layer(n, M, Q, Do, prec, tol)


# the (M, Q, Dout) of the BASELINE.json configs: cfg3 layers 1 / 2 (M=256, Q=2, Dout=2 / 1), cfg4
# (M=200, Q=4, Dout=4), cfg2 (M=128, Q=5, Dout=50), a Q=10 layer at M=256; n small enough for the
# oracle's [n,M,M] psi2, plus one n that spans several row tiles and row splits of the pair kernels
BENCH_MM = [(96, 256, 2, 2), (96, 256, 2, 1), (96, 200, 4, 4), (64, 128, 5, 50), (64, 256, 10, 1),
            (700, 256, 2, 2)]
BENCH_DET = [(256, 256, 10, 2), (300, 512, 16, 1), (256, 256, 10, 1)]


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-10), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,Q,Do', BENCH_MM)
def test_mm_layer_bench_shapes(n, M, Q, Do, prec, tol):
    """Every output of the moment-matched forward / backward ops against the oracle's
    psi_stats / psi_derivs at the benchmark's own pseudo-point counts (the wide DMMA kernels, the
    129-block pair grid at M = 256, row splits).  fp64 1e-10 (x10 on the cancelling sums) is three
    orders inside the 1e-6 bar."""
    oc.check_mm_layer(n, M, Q, Do, prec, tol)


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-10), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,D,Do', BENCH_DET)
def test_det_layer_bench_shapes(n, M, D, Do, prec, tol):
    oc.check_det_layer(n, M, D, Do, prec, tol)


def test_kmat_psi_lik():
    oc.check_kmat_psi_lik()


def test_torch_custom_ops():
    """`torch.ops.geepee_b200.*` (geepee_b200/torch_ops.py) on the device."""
    oc.check_torch_custom_ops()


def test_tail_primitives():
    """GpbTailOp program ops (batched DMMA GEMM, fused linear combinations, R packing, kernel-hyper
    chain rule ...) against numpy."""
    oc.check_tail_primitives()


@pytest.mark.parametrize('n,Do,Q', oc.EMIS_SHAPES)
def test_gauss_emis(n, Do, Q):
    oc.check_gauss_emis(n, Do, Q)


def test_probit_lik():
    oc.check_probit_lik()


@pytest.mark.parametrize('M,batch', [(5, 1), (32, 2), (50, 3), (77, 1), (128, 4), (200, 2), (256, 5), (512, 2)])
def test_spd_inverse(M, batch):
    oc.check_spd_inverse(M, batch)


# Row counts the oracle cannot reach: the fp32-psi kernels (tcgen05: several row tiles per persistent CTA, ring /
# accumulator phases carried from tile to tile, several fp64 flushes of the TMEM accumulators per row split) against
# the fp64 kernels of the same library, which the tests above pin to the oracle.
def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('n,M,D,Do', [(300000, 256, 10, 2), (150001, 200, 16, 1), (200000, 128, 3, 4)])
def test_det_layer_many_rows_fp32_vs_fp64(n, M, D, Do):
    from geepee_b200 import ops
    g = torch.Generator().manual_seed(n + M)
    dev = torch.device('cuda')

    def rnd(*s):
        return torch.randn(*s, generator=g, dtype=torch.float64).to(dev)
    x, z = rnd(n, D), rnd(M, D)
    ls = torch.full((D,), 0.4, dtype=torch.float64, device=dev)
    sf = torch.full((1,), 0.1, dtype=torch.float64, device=dev)
    A, B = rnd(Do, M), 0.05 * rnd(Do, M, M)
    B = (B + B.transpose(1, 2)).contiguous()
    dm, dv = rnd(n, Do), rnd(n, Do)
    out = {}
    for name in ('fp64', 'fp32'):
        pr = ops.PREC[name]
        opnd = ops.DetOperands(pr, A, B)
        m, v, Ks, Ts = ops.det_fwd(pr, x, z, ls, sf, opnd, save=True)
        dA, dzu, dl, dsf2 = ops.det_bwd(pr, x, z, ls, sf, opnd, dm, dv, Ks, Ts)
        dB = ops.det_syrk(pr, Ks, dv, M)
        out[name] = dict(m=m, v=v, dA=dA, dzu=dzu, dl=dl, dsf2=dsf2, dB=dB)
        del Ks, Ts
    for k in out['fp64']:
        assert _rel(out['fp32'][k], out['fp64'][k]) < 5e-4, (k, _rel(out['fp32'][k], out['fp64'][k]))


@pytest.mark.parametrize('n,M,Q,Do', [(100000, 128, 2, 2), (60000, 200, 4, 1), (50001, 64, 7, 3), (80000, 96, 3, 4)])
def test_mm_forward_many_rows_fp32_vs_fp64(n, M, Q, Do):
    from geepee_b200 import ops
    g = torch.Generator().manual_seed(n + M + Q)
    dev = torch.device('cuda')

    def rnd(*s):
        return torch.randn(*s, generator=g, dtype=torch.float64).to(dev)
    mx, z = rnd(n, Q), rnd(M, Q)
    vx = (0.05 + torch.rand(n, Q, generator=g, dtype=torch.float64)).to(dev)
    ls = torch.full((Q,), 0.3, dtype=torch.float64, device=dev)
    sf = torch.zeros(1, dtype=torch.float64, device=dev)
    A, B = rnd(Do, M), 0.05 * rnd(Do, M, M)
    B = (B + B.transpose(1, 2)).contiguous()
    o64 = ops.mm_fwd(ops.PREC['fp64'], mx, vx, z, ls, sf, A, B)
    o32 = ops.mm_fwd(ops.PREC['fp32'], mx, vx, z, ls, sf, A, B)
    for i, k in enumerate(('mout', 'vout', 'vacc', 'psi1')):
        assert _rel(o32[i], o64[i]) < 5e-4, (k, _rel(o32[i], o64[i]))


# edge shapes of the tcgen05 kernels: row counts around the 128-row tile, pseudo-point counts around the padding steps,
# every input-dimension template (4 / 8 / 16 / 32), several output dims (TMEM passes), feature counts around K = 8 / 16
@pytest.mark.parametrize('n,M,D,Do', [(1, 7, 1, 1), (127, 128, 4, 2), (129, 129, 5, 3), (255, 200, 8, 1), (257, 256, 9, 5),
                                       (1000, 31, 17, 2), (513, 256, 32, 4), (2049, 64, 3, 1)])
def test_det_layer_edge_shapes_fp32_vs_fp64(n, M, D, Do):
    test_det_layer_many_rows_fp32_vs_fp64(n, M, D, Do)


@pytest.mark.parametrize('n,M,Q,Do', [(1, 5, 1, 1), (127, 23, 3, 4), (129, 64, 4, 1), (255, 33, 7, 2), (257, 40, 2, 3),
                                       (640, 128, 1, 2), (385, 17, 6, 4)])
def test_mm_forward_edge_shapes_fp32_vs_fp64(n, M, Q, Do):
    test_mm_forward_many_rows_fp32_vs_fp64(n, M, Q, Do)
