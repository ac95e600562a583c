"""Pin the oracle (oracle/geepee_oracle.py + oracle/psi_oracle.c) against outputs of
the reference itself (tests/golden/*.npz, see tests/golden/gen_golden.py) and
against the reference's own identities.  CPU only."""
import copy

import numpy as np
import pytest

import golden_util as gu
import geepee_oracle as go

TOL = 1e-9   # observed <= 1e-12; the golden values come from LU-based np.linalg.inv


@pytest.mark.parametrize('name', gu.model_cases())
def test_models_match_reference(name):
    gold = gu.load(name)
    model = gu.build_oracle_model(gold)
    m = gold['meta']
    np.random.seed(m['rng_seed'])
    e, g = model.objective_function(copy.deepcopy(gold['p']), m['mb_size'], alpha=m['alpha'],
                                    prop_mode=m.get('prop_mode', 'MM'))
    gu.assert_close(e, g, gold, TOL, name)


def test_kernels_match_reference():
    f = np.load(gu.GOLDEN + '/kernels.npz')
    ls, sf, mx, vx, z = f['ls'], f['sf'], f['mx'], f['vx'], f['z']
    assert gu.rel_err(go.ard_kernel(2 * ls, 2 * sf, mx, z), f['kfu']) < 1e-14
    p1, p2 = go.psi_stats(2 * ls, 2 * sf, mx, vx, z)
    assert np.array_equal(p1, f['psi1']) and np.array_equal(p2, f['psi2'])   # same C loop: bit-exact
    q1, q2 = go.psi_stats_numpy(2 * ls, 2 * sf, mx, vx, z)
    assert gu.rel_err(q1, f['psi1_numpy']) < 1e-13 and gu.rel_err(q2, f['psi2_numpy']) < 1e-13
    assert gu.rel_err(q1, p1) < 1e-13 and gu.rel_err(q2, p2) < 1e-13
    d = go.psi_derivs(f['dpsi1'], p1, f['dpsi2'], p2, np.exp(ls), np.exp(2 * sf), mx, vx, z)
    for got, key in zip(d, ['var', 'l', 'z', 'mu', 'S']):
        assert gu.rel_err(got, f['psider_' + key]) < 1e-12, key
    d = go.kfu_derivs(f['dpsi1'], f['kfu'], np.exp(ls), np.exp(2 * sf), mx, z, grad_x=True)
    for got, key in zip(d, ['var', 'l', 'z', 'x']):
        assert gu.rel_err(got, f['kfuder_' + key]) < 1e-12, key
    t = go.dtrace_MKzz(2 * ls, 2 * sf, z, f['Mm'], f['Kzz'])
    for got, key in zip(t, ['sf', 'ls', 'z']):
        assert gu.rel_err(got, f['tr_' + key]) < 1e-12, key


def test_gauss_emis_matches_reference():
    f = np.load(gu.GOLDEN + '/gauss_emis.npz')
    em = go.GaussEmis(f['y'], f['y'].shape[1], f['mx'].shape[1])
    em.set_params({'C': f['C'], 'R': f['R']})
    idx = np.arange(f['y'].shape[0])
    lz, gi, gh = em.tilted(f['mx'], f['vx'], float(f['alpha']), float(f['scale']), idx)
    assert abs(lz - f['t_logZ']) < 1e-12 * abs(f['t_logZ'])
    for got, key in [(gi['mx'], 't_dmx'), (gi['vx'], 't_dvx'), (gh['C'], 't_dC'), (gh['R'], 't_dR')]:
        assert gu.rel_err(got, f[key]) < 1e-12, key
    le, gi, gh = em.log_lik_exp(f['mx'], f['vx'], float(f['scale']), idx)
    assert abs(le - f['e_logZ']) < 1e-12 * abs(f['e_logZ'])
    for got, key in [(gi['mx'], 'e_dmx'), (gi['vx'], 'e_dvx'), (gh['C'], 'e_dC'), (gh['R'], 'e_dR')]:
        assert gu.rel_err(got, f[key]) < 1e-12, key


def test_psi_with_zero_variance_is_the_kernel():
    """TODO.txt:65-66 / SURVEY 8c(iv): psi1(vx=0) == kfu, psi2(vx=0) == kfu (x) kfu."""
    rng = np.random.RandomState(0)
    mx, z = rng.standard_normal((5, 3)), rng.standard_normal((4, 3))
    ls, sf = 0.2 * rng.standard_normal(3), np.array([0.1])
    k = go.ard_kernel(2 * ls, 2 * sf, mx, z)
    p1, p2 = go.psi_stats(2 * ls, 2 * sf, mx, np.zeros_like(mx), z)
    assert gu.rel_err(p1, k) < 1e-14
    assert gu.rel_err(p2, k[:, :, None] * k[:, None, :]) < 1e-14


def test_aep_alpha_to_zero_is_vfe():
    """tests/test_aep_vfe_limits.py:17-34 (SGPR): AEP energy at alpha=1e-6 equals the VFE
    energy at the same parameters (needs alpha*vout/sn2 << 1: SURVEY A6.1)."""
    gold = gu.load('aep_sgpr')
    i, m = gold['in'], gold['meta']
    p = copy.deepcopy(gold['p'])
    p['sn'] = np.array(np.log(0.5))
    ea, _ = go.AepSGPR(i['x'], i['y'], m['M']).objective_function(copy.deepcopy(p), m['N'], alpha=1e-6)
    ev, _ = go.VfeSGPR(i['x'], i['y'], m['M']).objective_function(copy.deepcopy(p), m['N'])
    assert abs(float(np.ravel(ea)[0]) - float(np.ravel(ev)[0])) < 1e-4 * abs(float(np.ravel(ev)[0]))


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr', 'aep_sgplvm', 'aep_sgpssm_lin_1d', 'vfe_sgpr'])
def test_finite_differences(name):
    """The reference's own harness (tests/test_utils.py:61-138): central differences,
    eps=1e-5, pass if rel diff < 1e-4 (or both tiny).  A few random entries per key."""
    gold = gu.load(name)
    model = gu.build_oracle_model(gold)
    m = gold['meta']
    N, alpha = m['N'], m['alpha']
    p0 = gold['p']
    _, g = model.objective_function(copy.deepcopy(p0), N, alpha=alpha)
    rng = np.random.RandomState(1)
    eps = 1e-5
    for key in sorted(p0):
        flat = np.asarray(p0[key]).reshape(-1)
        for j in rng.choice(flat.size, size=min(3, flat.size), replace=False):
            vals = []
            for sgn in (+1, -1):
                p = copy.deepcopy(p0)
                q = np.array(p[key], dtype=np.float64)
                q.reshape(-1)[j] += sgn * eps
                p[key] = q
                e, _ = model.objective_function(p, N, alpha=alpha)
                vals.append(float(np.ravel(e)[0]))
            num = (vals[0] - vals[1]) / (2 * eps)
            ana = float(np.asarray(g[key]).reshape(-1)[j])
            assert abs(ana - num) <= 2e-4 * max(abs(num), abs(ana)) + 1e-6, (name, key, j, ana, num)
