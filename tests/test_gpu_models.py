"""Model-level parity on the B200: geepee_b200 models (reference API) through the C ABI against
the golden vectors generated from the reference -- energy and EVERY gradient key, 1e-6 relative
in fp64 mode and 1e-3 in fp32-psi mode (BASELINE.json north_star) -- plus size-independent
properties at larger shapes where the oracle would take too long."""
import copy

import numpy as np
import pytest
import torch

import golden_util as gu
import model_cases as mc

pytestmark = pytest.mark.gpu

TOL64 = 1e-6
TOL32 = 1e-3


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()
    yield


@pytest.mark.parametrize('name', gu.model_cases())
def test_objective_fp64(name):
    mc.check_model(name, 'fp64', TOL64)


@pytest.mark.parametrize('name', gu.model_cases())
def test_objective_fp32(name):
    gold = gu.load(name)
    if max(gold['meta'].get('floor', {}).values() or [0]) > 1e-7:
        pytest.skip('ill-conditioned case (reference floor > 1e-7): fp64 only')
    mc.check_model(name, 'fp32', TOL32)


@pytest.mark.parametrize('prec,tol', [('fp64', TOL64), ('fp32', TOL32)])
@pytest.mark.parametrize('name', gu.model_cases(bench=True))
def test_objective_bench_shapes(name, prec, tol):
    """Reduced-n runs of the BASELINE.json configs at their own M / Q / Dout (M = 256, 200, 128, 512),
    against the reference itself (tests/golden/gen_golden_bench.py), both precisions."""
    gold = gu.load(name)
    if prec == 'fp32' and max(gold['meta'].get('floor', {}).values() or [0]) > 1e-7:
        pytest.skip('ill-conditioned parameter point (the reference itself moves by > 1e-7 under a 1e-15 '
                    'perturbation, meta.floor): fp64 only -- the well-conditioned twin *_wc covers fp32')
    mc.check_model(name, prec, tol)


@pytest.mark.parametrize('name', ['aep_sgpr', 'vfe_sgpr', 'aep_sdgpr', 'aep_sgpr_nonnat', 'aep_sgpr_cfg1', 'aep_sdgprh'])
def test_predict(name):
    mc.check_predict(name, 'fp64', 1e-7)


@pytest.mark.parametrize('name', ['aep_sgpr', 'vfe_sgpr', 'aep_sgpr_probit', 'bench_ns_sgpr_n2048'])
def test_chunked_rows(name):
    """config.DET_SAVE_BYTES: the single-layer models process rows in chunks whose saved Kfu / T tiles stay
    bounded; forcing many chunks reproduces the reference."""
    mc.check_chunked(name, 'fp64', TOL64)


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr'])
def test_sampling(name):
    mc.check_sampling(name, 1e-5)


@pytest.mark.parametrize('name', ['aep_sgpssm_lin', 'aep_sgpssm_gp'])
def test_ssm_predict(name):
    mc.check_ssm_predict(name, 1e-7)


def _oracle_vs_gpu(make_oracle, make_gpu, params, N, alpha, tol):
    """Compare with the oracle on the same inputs.  The oracle's own conditioning floor (how far
    its output moves when every input is perturbed by 1e-15 relative, as in gen_golden.py) bounds
    what any implementation can match; the tolerance per key is max(tol, 10*floor)."""
    om = make_oracle()
    eo, g_o = om.objective_function(copy.deepcopy(params), N, alpha=alpha)
    rng = np.random.RandomState(999)
    floor = {k: 0.0 for k in g_o}
    for _ in range(2):
        q = {k: np.array(v, dtype=np.float64) * (1.0 + 1e-15 * rng.standard_normal(np.shape(v)))
             for k, v in params.items()}
        _, g2 = om.objective_function(q, N, alpha=alpha)
        for k in g_o:
            floor[k] = max(floor[k], gu.rel_err(g2[k], g_o[k]))
    eg, g_g = make_gpu().objective_function(copy.deepcopy(params), N, alpha=alpha)
    gold = {'energy': float(np.ravel(eo)[0]), 'g': g_o, 'meta': {'floor': floor}}
    gu.assert_close(eg, g_g, gold, tol, 'oracle-vs-gpu')


def test_sgpr_medium_vs_oracle():
    """N=3000, M=60, D=4: several tiles / blocks / row splits, compared with the oracle."""
    import geepee_oracle as go
    from geepee_b200 import aep_models as aep
    rng = np.random.RandomState(0)
    N, M, D, Do = 3000, 60, 4, 2
    x = rng.standard_normal((N, D))
    y = np.sin(x[:, :Do]) + 0.1 * rng.standard_normal((N, Do))
    np.random.seed(0)
    model = aep.SGPR(x, y, M)
    p = model.init_hypers(y)
    p['sn'] = np.array(np.log(0.2))
    _oracle_vs_gpu(lambda: go.AepSGPR(x, y, M), lambda: model, p, N, 0.5, TOL64)


def test_sdgpr_medium_vs_oracle():
    import geepee_oracle as go
    from geepee_b200 import aep_models as aep
    rng = np.random.RandomState(1)
    N, M, D = 400, 30, 3
    x = rng.standard_normal((N, D))
    y = np.sin(x[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    np.random.seed(1)
    model = aep.SDGPR(x, y, M, [2, 2])
    p = model.init_hypers(y)
    p['sn'] = np.array(np.log(0.2))
    _oracle_vs_gpu(lambda: go.AepSDGPR(x, y, M, [2, 2]), lambda: model, p, N, 1.0, TOL64)


def test_linearity_in_rows_large():
    """Size-independent property at a large shape: the per-row statistics are additive over
    rows, so energy*N - phi over [rows A + rows B] equals the sum of the two halves.  Checked
    through the public objective by comparing a full batch to the mean of its two half-batch
    'datasets' is not exact for AEP (N enters phi), so check the additive kernels directly."""
    from geepee_b200 import ops
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(0)
    n, M, D, Do = 200000, 256, 10, 1
    x = torch.randn(n, D, generator=g, dtype=torch.float64).to(dev)
    z = torch.randn(M, D, generator=g, dtype=torch.float64).to(dev)
    ls = torch.full((D,), 0.7, dtype=torch.float64, device=dev)
    sf = torch.zeros(1, dtype=torch.float64, device=dev)
    A = torch.randn(Do, M, generator=g, dtype=torch.float64).to(dev)
    B = 0.01 * torch.randn(Do, M, M, generator=g, dtype=torch.float64).to(dev)
    B = (B + B.transpose(1, 2)).contiguous()
    dm = torch.randn(n, Do, generator=g, dtype=torch.float64).to(dev)
    dv = torch.randn(n, Do, generator=g, dtype=torch.float64).to(dev)
    opnd = ops.DetOperands(ops.F64, A, B)

    def stats(lo, hi):
        xs, dms, dvs = x[lo:hi].contiguous(), dm[lo:hi].contiguous(), dv[lo:hi].contiguous()
        m, v, Ks, Ts = ops.det_fwd(ops.F64, xs, z, ls, sf, opnd, save=True)
        dA, dzu, dl, dsf2 = ops.det_bwd(ops.F64, xs, z, ls, sf, opnd, dms, dvs, Ks, Ts)
        return [dA, dzu, dl, dsf2, ops.det_syrk(ops.F64, Ks, dvs, M), m.sum(0), v.sum(0)]

    full = stats(0, n)
    a, b = stats(0, 70001), stats(70001, n)
    for f, p, q in zip(full, a, b):
        assert gu.rel_err((p + q).cpu().numpy(), f.cpu().numpy()) < 1e-10


def test_sgplvm_wide_medium_vs_oracle():
    """SGPLVM with a wide output layer (Do = 20 -> the row-owner forward and the 32-dims-per-pass
    backward), several pair chunks and row splits, against the oracle."""
    import geepee_oracle as go
    import bench
    from geepee_b200 import aep_models as aep
    w = dict(model='SGPLVM', N=600, Q=3, Do=20, M=40, alpha=0.5, seed=2)
    X, Y = bench.make_data(w)
    p = bench.make_params(None, Y, w, X)
    p['ls'] = 0.3 * np.ones(w['Q'])
    _oracle_vs_gpu(lambda: go.AepSGPLVM(Y, w['Q'], w['M']), lambda: aep.SGPLVM(Y, w['Q'], w['M']),
                   p, w['N'], w['alpha'], TOL64)


def test_sgpssm_medium_vs_oracle():
    """SGPSSM with the fused emission kernel and the Q = 4 pair kernels (4 pairs per thread)."""
    import geepee_oracle as go
    import bench
    from geepee_b200 import aep_models as aep
    w = dict(model='SGPSSM', N=500, Q=4, Do=4, M=30, alpha=0.5, seed=4)
    X, Y = bench.make_data(w)
    p = bench.make_params(None, Y, w, X)
    _oracle_vs_gpu(lambda: go.AepSGPSSM(Y, w['Q'], w['M']), lambda: aep.SGPSSM(Y, w['Q'], w['M']),
                   p, w['N'], w['alpha'], TOL64)


def test_pair_kernels_additive_in_rows_large():
    """Size-independent property of the moment-matched layer at a large shape (n = 100 000, M = 128):
    the statistics of a batch equal the sum over two row blocks."""
    from geepee_b200 import ops
    dev = torch.device('cuda')
    g = torch.Generator().manual_seed(1)
    n, M, Q, Do = 100000, 128, 3, 2
    mx = torch.randn(n, Q, generator=g, dtype=torch.float64).to(dev)
    vx = (0.05 + torch.rand(n, Q, generator=g, dtype=torch.float64)).to(dev)
    z = torch.randn(M, Q, generator=g, dtype=torch.float64).to(dev)
    ls = torch.full((Q,), 0.2, dtype=torch.float64, device=dev)
    sf = torch.zeros(1, dtype=torch.float64, device=dev)
    A = torch.randn(Do, M, generator=g, dtype=torch.float64).to(dev)
    B = (0.01 * torch.randn(Do, M, M, generator=g, dtype=torch.float64)).to(dev).contiguous()
    dm = torch.randn(n, Do, generator=g, dtype=torch.float64).to(dev)
    dv = torch.randn(n, Do, generator=g, dtype=torch.float64).to(dev)

    def stats(lo, hi):
        a, b, c, d = (t[lo:hi].contiguous() for t in (mx, vx, dm, dv))
        mo, vo, va, p1 = ops.mm_fwd(ops.F64, a, b, z, ls, sf, A, B)
        o = ops.mm_bwd(ops.F64, a, b, z, ls, sf, A, B, c, d, mo, va, p1)
        return [o['dA'], o['dB'], o['dzu'], o['dl'], o['dsf2'], o['dvsum'], mo.sum(0), vo.sum(0),
                o['dmx'].sum(0), o['dvx'].sum(0)]

    full = stats(0, n)
    a, b = stats(0, 33333), stats(33333, n)
    for f, p, q in zip(full, a, b):
        assert gu.rel_err((p + q).cpu().numpy(), f.cpu().numpy()) < 1e-9
