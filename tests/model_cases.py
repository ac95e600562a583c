"""Shared bodies of the model-level parity checks: geepee_b200 models (the reference's API)
against the golden vectors generated from the reference (tests/golden).  Run on the CPU fiber
emulator by tests/test_emu_models.py and on the B200 by tests/test_gpu_models.py."""
import copy

import numpy as np

import golden_util as gu


def build_model(gold, prec='fp64', device=None):
    from geepee_b200 import aep_models as aep, vfe_models as vfe
    m, i = gold['meta'], gold['in']
    kind = m['model']
    kw = dict(prec=prec, device=device)
    lk = m.get('lik', 'Gaussian')
    if kind == 'aep_models.SGPR':
        return aep.SGPR(i['x'], i['y'], m['M'], lik=lk, nat_param=m['nat_param'], **kw)
    if kind == 'vfe_models.SGPR':
        return vfe.SGPR(i['x'], i['y'], m['M'], lik=lk, nat_param=m['nat_param'], **kw)
    if kind == 'aep_models.SDGPR':
        return aep.SDGPR(i['x'], i['y'], m['M'], m['hidden'], lik=lk, **kw)
    if kind == 'aep_models.SDGPR_H':
        return aep.SDGPR_H(i['x'], i['y'], m['M'], m['hidden'], lik=lk, **kw)
    if kind == 'aep_models.SGPLVM':
        return aep.SGPLVM(i['y'], m['Q'], m['M'], lik=lk, **kw)
    if kind == 'vfe_models.SGPLVM':
        return vfe.SGPLVM(i['y'], m['Q'], m['M'], lik=lk, nat_param=m['nat_param'], **kw)
    if kind == 'aep_models.SGPSSM':
        return aep.SGPSSM(i['y'], m['Q'], m['M'], x_control=i.get('x_control'), gp_emi=m['gp_emi'], **kw)
    if kind == 'vfe_models.SGPSSM':
        return vfe.SGPSSM(i['y'], m['Q'], m['M'], x_control=i.get('x_control'), gp_emi=m['gp_emi'],
                          nat_param=m['nat_param'], **kw)
    raise ValueError(kind)


def check_model(name, prec, tol, device=None):
    gold = gu.load(name)
    model = build_model(gold, prec, device)
    m = gold['meta']
    np.random.seed(m['rng_seed'])
    e, g = model.objective_function(copy.deepcopy(gold['p']), m['mb_size'], alpha=m['alpha'],
                                    prop_mode=m.get('prop_mode', 'MM'))
    gu.assert_close(e, g, gold, tol, '%s[%s]' % (name, prec))
    return model, gold


def check_chunked(name, prec, tol, device=None):
    """Row-chunked deterministic step (config.DET_SAVE_BYTES): forcing several chunks must reproduce the
    golden energy and gradients (the statistics are additive over rows)."""
    from geepee_b200 import config
    old = config.DET_SAVE_BYTES, config.DET_MIN_CHUNK_ROWS
    config.DET_SAVE_BYTES, config.DET_MIN_CHUNK_ROWS = 1, 7
    try:
        model, gold = check_model(name, prec, tol, device)
        assert len(model.sgp_layer.det_chunks(gold['meta']['mb_size'])) > 1
    finally:
        config.DET_SAVE_BYTES, config.DET_MIN_CHUNK_ROWS = old


def check_predict(name, prec, tol, device=None):
    gold = gu.load(name)
    model = build_model(gold, prec, device)
    model.update_hypers(copy.deepcopy(gold['p']))
    model.updated = False
    x = gold['x']
    mf, vf = model.predict_f(x['xs'])
    my, vy = model.predict_y(x['xs'])
    for got, key in [(mf, 'mf'), (vf, 'vf'), (my, 'my'), (vy, 'vy')]:
        assert gu.rel_err(got, x[key]) < tol, (name, key, gu.rel_err(got, x[key]))


def check_ssm_predict(name, tol, device=None):
    """Prediction API of the state-space model around the hot path (base_models.py:1453-1595):
    predict_forward_mm / _mc, predict_y, get_posterior_y against the reference's outputs."""
    gold = gu.load(name)
    model = build_model(gold, 'fp64', device)
    model.update_hypers(copy.deepcopy(gold['p']))
    x = gold['x']
    xc = x.get('pf_xc')
    pf = model.predict_forward_mm(x['pf_mx'].shape[0], xc)
    for got, key in zip(pf, ('pf_mx', 'pf_vx', 'pf_my', 'pf_vyn', 'pf_vy')):
        assert gu.rel_err(got, x[key]) < tol, (name, key, gu.rel_err(got, x[key]))
    py = model.predict_y(x['py_in'])
    for got, key in zip(py, ('py_my', 'py_vy')):
        assert gu.rel_err(got, x[key]) < tol, (name, key, gu.rel_err(got, x[key]))
    gy = model.get_posterior_y()
    for got, key in zip(gy, ('gy_my', 'gy_vf', 'gy_vyn')):
        assert gu.rel_err(got, x[key]) < tol, (name, key, gu.rel_err(got, x[key]))
    if 'mc_x' in x:
        np.random.seed(321)
        T, S = x['mc_x'].shape[0], x['mc_x'].shape[1]
        pmc = model.predict_forward(T, xc, prop_mode='MC', no_samples=S)
        for got, key in zip(pmc, ('mc_x', 'mc_my', 'mc_vy')):
            assert gu.rel_err(got, x[key]) < tol, (name, key, gu.rel_err(got, x[key]))


def check_sampling(name, tol, device=None):
    """sample_f (base_models.py:428-452, 1000-1018, 1239-1262) and the particle prediction of the
    deep GP (1160-1184), seeded like the golden generator.  The f | u draw factorises
    kff - kfu Kuu^-1 kuf, which is close to singular by construction: tolerance 1e-5."""
    gold = gu.load(name)
    model = build_model(gold, 'fp64', device)
    model.update_hypers(copy.deepcopy(gold['p']))
    model.updated = False
    x = gold['x']
    if 'mc_samples' in x:
        np.random.seed(556)
        smp, mf, vf = model.predict_f(x['xs'], prop_mode='MC', no_samples=x['mc_samples'].shape[0])
        for got, key in ((smp, 'mc_samples'), (mf, 'mc_mf'), (vf, 'mc_vf')):
            assert gu.rel_err(got, x[key]) < 1e-7, (name, key, gu.rel_err(got, x[key]))
        np.random.seed(557)
    else:
        np.random.seed(555)
    fs = model.sample_f(x['xs'], x['fs'].shape[2])
    assert gu.rel_err(fs, x['fs']) < tol, (name, 'fs', gu.rel_err(fs, x['fs']))


def check_input_grad(tol, device=None):
    """predict_f_with_input_grad / predict_y_with_input_grad / layer.backprop_predictive_grads_reg
    (base_models.py:1186-1237,1277-1289,391-426) against outputs of the reference itself
    (tests/golden/input_grad.npz, gen_golden.py::case_input_grad)."""
    import json
    import os
    from geepee_b200 import aep_models as aep
    f = np.load(os.path.join(gu.GOLDEN, 'input_grad.npz'), allow_pickle=False)
    for c in json.loads(str(f['meta']))['cases']:
        t = c['tag'] + '_'
        g = {k[len(t):]: np.array(f[k]) for k in f.files if k.startswith(t)}
        params = {k[2:]: v for k, v in g.items() if k.startswith('p_')}
        model = aep.SDGPR(g['x'], g['y'], c['M'], [], lik='Gaussian', prec='fp64', device=device)
        model.update_hypers(copy.deepcopy(params))
        model.updated = False
        mf, vf, dm_dx, dv_dx = model.predict_f_with_input_grad(g['xs'])
        my, vy, dm2, dv2 = model.predict_y_with_input_grad(g['xs'])
        layer = model.sgp_layers[0]
        m0, v0, kfu = layer.forward_prop_thru_post(g['xs'], return_info=True)
        l_dm, l_dv = layer.backprop_predictive_grads_reg(m0, v0, g['w_m'], g['w_v'], np.zeros((1, 1)),
                                                         np.ones((1, 1)), kfu, g['xs'])
        for got, key in ((mf, 'mf'), (vf, 'vf'), (dm_dx, 'dm_dx'), (dv_dx, 'dv_dx'), (my, 'my'), (vy, 'vy'),
                         (dm2, 'dm_dx'), (dv2, 'dv_dx'), (l_dm, 'l_dm'), (l_dv, 'l_dv')):
            assert got.shape == g[key].shape, (c, key, got.shape, g[key].shape)
            assert gu.rel_err(got, g[key]) < tol, (c, key, gu.rel_err(got, g[key]))
    # deeper models: the reference fails on the undefined lvm_mm twin (base_models.py:1231)
    deep = aep.SDGPR(g['x'], g['y'], c['M'], [2], lik='Gaussian', prec='fp64', device=device)
    deep.update_hypers(deep.init_hypers(g['y']))
    try:
        deep.predict_f_with_input_grad(g['xs'])
    except AttributeError:
        return
    raise AssertionError('predict_f_with_input_grad on a 2-layer model should fail as the reference does')


def check_layer_iface(tol, device=None):
    """SURVEY 8b layer-level interface (compute_cavity / forward_prop_thru_cav / backprop_grads_reg /
    backprop_grads_lvm_mm / compute_phi* / compute_KL / forward_prop_thru_post) against outputs of the
    reference's layers (tests/golden/layer_iface.npz, gen_golden.py::case_layer_iface)."""
    import json
    import os
    from geepee_b200 import aep_models as aep, vfe_models as vfe
    f = np.load(os.path.join(gu.GOLDEN, 'layer_iface.npz'), allow_pickle=False)
    for c in json.loads(str(f['meta']))['cases']:
        t = c['tag'] + '__'
        g = {k[len(t):]: np.array(f[k]) for k in f.files if k.startswith(t)}
        params = {k[2:]: v for k, v in g.items() if k.startswith('p_')}
        cls = aep.SGP_Layer if c['mod'] == 'aep' else vfe.SGP_Layer
        layer = cls(c['N'], c['D'], c['Do'], c['M'], c['nat'], 'fp64', device)
        layer.update_hypers(copy.deepcopy(params))
        layer.compute_kuu()
        layer.update_posterior()
        alpha = c['alpha']
        x, mx, vx = g['x'], g['mx'], g['vx']
        if c['mod'] == 'aep':
            layer.compute_cavity(alpha)
            phi = np.array([layer.compute_phi(alpha), layer.compute_phi_prior(),
                            layer.compute_phi_posterior(), layer.compute_phi_cavity()])
            m, v, kfu = layer.forward_prop_thru_cav(x)
            gr = layer.backprop_grads_reg(m, v, g['dm'], g['dv'], kfu, x, alpha)
            ms, vs, psi1, psi2 = layer.forward_prop_thru_cav(mx, vx, mode='MM')
            if c['nat']:
                gs, gx = layer.backprop_grads_lvm_mm(ms, vs, g['dm2'], g['dv2'], psi1, psi2, mx, vx, alpha)
            else:
                # the reference applies its natural-parameter chain rule here whatever nat_param is
                # (aep_models.py:252-297 use theta_2 / theta_1_R unconditionally); the product refuses
                # instead of reproducing those numbers, so the recorded gs_* / gx_* are not compared
                try:
                    layer.backprop_grads_lvm_mm(ms, vs, g['dm2'], g['dv2'], psi1, psi2, mx, vx, alpha)
                except NotImplementedError:
                    gs, gx = None, None
                else:
                    raise AssertionError('non-natural AEP moment-matched backprop should refuse')
        else:
            phi = np.array([layer.compute_KL()])
            m, v, kfu = layer.forward_prop_thru_post(x, return_info=True)
            gr = layer.backprop_grads_reg(m, v, g['dm'], g['dv'], kfu, x)
            ms, vs, psi1, psi2 = layer.forward_prop_thru_post(mx, vx, mode='MM', return_info=True)
            gs, gx = layer.backprop_grads_lvm_mm(ms, vs, g['dm2'], g['dv2'], psi1, psi2, mx, vx)
        pm, pv = layer.forward_prop_thru_post(x)
        pms, pvs = layer.forward_prop_thru_post(mx, vx, mode='MM')
        got = dict(phi=phi, m=m, v=v, kfu=kfu, ms=ms, vs=vs, psi1=psi1, psi2=psi2, pm=pm, pv=pv, pms=pms,
                   pvs=pvs)
        if 'mc_m' in g:         # Monte-Carlo propagation at layer level, eps from the same numpy seed
            np.random.seed(321)
            if c['mod'] == 'aep':
                res, res_s = layer.forward_prop_thru_cav(mx, vx, mode='MC')
                gmc, dxs = layer.backprop_grads_lvm_mc(res_s[0], res_s[1], g['dm3'], g['dv3'], res_s[2],
                                                       res_s[3], alpha)
            else:
                res, res_s = layer.forward_prop_thru_post(mx, vx, mode='MC', return_info=True)
                gmc, dxs = layer.backprop_grads_lvm_mc(res_s[0], res_s[1], g['dm3'], g['dv3'], res_s[2], res_s[3])
            gin = layer.backprop_grads_reparam(dxs, mx, vx, res[4])
            K, n = res[0].shape[:2]
            for a, b in zip(res, res_s):
                assert np.array_equal(a.reshape(b.shape), b)
            got.update(mc_m=res[0], mc_v=res[1], mc_kfu=res[2], mc_x=res[3], mc_eps=res[4], mc_dxs=dxs,
                       mc_gx_mx=gin['mx'], mc_gx_vx=gin['vx'])
            got.update({'mcg_' + k: a for k, a in gmc.items()})
            np.random.seed(322)
            got['mc_pm'], got['mc_pv'] = layer.forward_prop_thru_post(mx, vx, mode='MC')
        got.update({'g_' + k: a for k, a in gr.items()})
        want = {k for k in g if not k.startswith('p_')} - {'xtr', 'x', 'mx', 'vx', 'dm', 'dv', 'dm2', 'dv2', 'dm3', 'dv3'}
        if gs is not None:
            got.update(gx_mx=gx['mx'], gx_vx=gx['vx'])
            got.update({'gs_' + k: a for k, a in gs.items()})
        else:
            want = {k for k in want if not k.startswith(('gs_', 'gx_'))}
        assert set(got) == want, (c['tag'], sorted(set(got) ^ want))
        for k in sorted(want):
            a = np.asarray(got[k], dtype=np.float64)
            assert a.size == g[k].size, (c['tag'], k, a.shape, g[k].shape)
            if k == 'phi':      # four independent scalars
                for i in range(a.size):
                    assert abs(a[i] - g[k][i]) <= tol * max(abs(g[k][i]), 1e-12), (c['tag'], k, i, a, g[k])
                continue
            assert gu.rel_err(a, g[k]) < tol, (c['tag'], k, gu.rel_err(a, g[k]))


def check_aep_to_vfe_limit(name, device=None, alpha=1e-6, tol=1e-5):
    """tests/test_aep_vfe_limits.py:17-127 through the product: the AEP energy at alpha -> 0 equals
    the VFE energy at the same parameters (aep.SGPLVM is not divided by N: aep_models.py:803-815)."""
    gold = gu.load(name)
    m = gold['meta']
    p = copy.deepcopy(gold['p'])
    if 'sn' in p:
        p['sn'] = np.array(np.log(0.5))       # alpha*vout/sn2 << 1
    ev, _ = build_model(gold, 'fp64', device).objective_function(copy.deepcopy(p), m['N'])
    g2 = dict(gold)
    g2['meta'] = dict(m, model=m['model'].replace('vfe_', 'aep_'))
    ea, _ = build_model(g2, 'fp64', device).objective_function(copy.deepcopy(p), m['N'], alpha=alpha)
    ev, ea = float(np.ravel(ev)[0]), float(np.ravel(ea)[0])
    if 'SGPLVM' in m['model']:
        ea /= m['N']
    assert abs(ea - ev) < tol * abs(ev), (name, ea, ev)


def check_finite_differences(name, device=None, per_key=2):
    """The reference's own gradient harness (tests/test_utils.py:61-138, used by tests/test_grads_*.py)
    applied to the product: central differences with eps = 1e-5 on a few random entries of every
    parameter key against the analytic gradient the kernels return."""
    gold = gu.load(name)
    model = build_model(gold, 'fp64', device)
    m = gold['meta']
    mb, alpha = m['N'], m['alpha']
    p0 = gold['p']
    _, g = model.objective_function(copy.deepcopy(p0), mb, alpha=alpha)
    rng = np.random.RandomState(1)
    eps = 1e-5
    for key in sorted(p0):
        flat = np.asarray(p0[key]).reshape(-1)
        for j in rng.choice(flat.size, size=min(per_key, flat.size), replace=False):
            vals = []
            for sgn in (+1, -1):
                p = copy.deepcopy(p0)
                q = np.array(p[key], dtype=np.float64)
                q.reshape(-1)[j] += sgn * eps
                p[key] = q
                e, _ = model.objective_function(p, mb, alpha=alpha)
                vals.append(float(np.ravel(e)[0]))
            num = (vals[0] - vals[1]) / (2 * eps)
            ana = float(np.asarray(g[key]).reshape(-1)[j])
            assert abs(ana - num) <= 2e-4 * max(abs(num), abs(ana)) + 1e-6, (name, key, j, ana, num)


def check_lik_iface(tol, device=None):
    """Public likelihood-layer interface, 2-D and Monte-Carlo 3-D branches, against outputs of the
    reference's Gauss_Layer / Probit_Layer (tests/golden/lik_iface.npz, gen_golden.py::case_lik_iface)."""
    import json
    import os
    from geepee_b200 import lik_layers as plik
    f = np.load(os.path.join(gu.GOLDEN, 'lik_iface.npz'), allow_pickle=False)
    meta = json.loads(str(f['meta']))
    for c in meta['cases']:
        t = c['tag'] + '__'
        g = {k[len(t):]: np.array(f[k]) for k in f.files if k.startswith(t)}
        L = getattr(plik, c['lik'] + '_Layer')(c['n'], c['D'], device)
        if c['lik'] == 'Gauss':
            L.update_hypers({'sn': np.array(meta['sn'])})
        m, v, y, alpha = g['m'], g['v'], g['y'], c['alpha']
        v1 = v.copy()
        r = L.compute_log_Z(m, v1, y, alpha)
        gr = L.backprop_grads(m, v1, r[1], r[2], alpha, 0.7)
        e = L.compute_log_lik_exp(m, v, y)
        ge = L.backprop_grads_log_lik_exp(m, v, e[1], e[2], y, 0.7)
        got = dict(logZ=[r[0]], dm=r[1], dv=r[2], ll=[e[0]], edm=e[1], edv=e[2])
        if c['lik'] == 'Gauss':
            got.update(vout=v1, g_sn=[gr['sn']], ge_sn=[ge['sn']])
            if m.ndim == 2:
                got['dm2'] = L.compute_log_Z(m, v.copy(), y, alpha, compute_dm2=True)[3]
        else:
            assert gr == {} and ge == {}
            assert np.array_equal(v1, v)          # the probit layer leaves vout alone
        for k, a in got.items():
            a = np.asarray(a, dtype=np.float64)
            assert a.shape == g[k].shape, (c['tag'], k, a.shape, g[k].shape)
            assert gu.rel_err(a, g[k]) < tol, (c['tag'], k, gu.rel_err(a, g[k]))
    L = plik.Gauss_Layer(3, 2, device)
    L.update_hypers({'sn': np.array(0.0)})
    try:
        L.compute_log_Z(np.zeros(3), np.ones(3), np.zeros(3))
    except RuntimeError:
        return
    raise AssertionError('1-D input should raise RuntimeError (lik_layers.py:152)')


def check_gauss_emis_limit(device=None):
    """tests/test_grads_emis.py:212-240 through the product: the tilted emission log-partition at
    alpha -> 0, scaled by 1/alpha, equals the expected log-likelihood (SURVEY 8c pin iii)."""
    from geepee_b200 import lik_layers as plik
    rng = np.random.RandomState(11)
    N, Dout, Din, alpha = 5, 3, 2, 1e-5
    y = rng.standard_normal((N, Dout))
    emis = plik.Gauss_Emis(y, Dout, Din, device)
    emis.update_hypers({'C': rng.standard_normal((Dout, Din)), 'R': rng.standard_normal(Dout)})
    mx, vx = rng.standard_normal((N, Din)), rng.rand(N, Din)
    z1, gi1, gp1 = emis.compute_emission_tilted(mx, vx, alpha, 1.0 / alpha)
    z2, gi2, gp2 = emis.compute_emission_log_lik_exp(mx, vx, 1.0)
    assert abs(z1 - z2) < 1e-3 * abs(z2), (z1, z2)
    for a, b in ((gi1, gi2), (gp1, gp2)):
        assert set(a) == set(b)
        for k in a:
            assert gu.rel_err(a[k], b[k]) < 1e-3, (k, gu.rel_err(a[k], b[k]))


def check_psi_zero_variance_is_kernel(device=None):
    """TODO.txt:65-66 / SURVEY 8c pin (iv) through the product: psi1(vx = 0) = Kfu and
    psi2(vx = 0) = Kfu (x) Kfu, and the moment-matched layer with vx = 0 returns the deterministic
    layer's mean (its variance differs by the Bhat_sto / Bhat_det choice, aep_models.py:155,196)."""
    import torch
    from geepee_b200 import ops
    from geepee_b200.layers import to_dev, default_device
    dev = device or default_device()
    rng = np.random.RandomState(0)
    mx, z = to_dev(rng.standard_normal((9, 3)), dev), to_dev(rng.standard_normal((6, 3)), dev)
    ls, sf = to_dev(0.2 * rng.standard_normal(3), dev), to_dev(np.array([0.1]), dev)
    k = ops.kmat(mx, z, ls, sf).cpu().numpy()
    p1, p2 = ops.psi_stats(mx, torch.zeros_like(mx), z, ls, sf)
    assert gu.rel_err(p1.cpu().numpy(), k) < 1e-13
    assert gu.rel_err(p2.cpu().numpy(), k[:, :, None] * k[:, None, :]) < 1e-13
    gold = gu.load('aep_sgplvm')
    model = build_model(gold, 'fp64', device)
    model.update_hypers(copy.deepcopy(gold['p']))
    layer = model.sgp_layer
    layer.compute_cavity(0.5)
    x = rng.standard_normal((7, layer.Din))
    md, _, _ = layer.forward_prop_thru_cav(x)
    ms, _, _, _ = layer.forward_prop_thru_cav(x, np.zeros_like(x), mode='MM')
    assert gu.rel_err(ms, md) < 1e-9, gu.rel_err(ms, md)
