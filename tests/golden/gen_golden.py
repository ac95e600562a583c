#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container (needs /root/reference):
    python tests/golden/gen_golden.py

It imports the mechanically py3-patched copy of the reference that
oracle/make_ref.py builds (git-ignored oracle/_ref/), evaluates
``objective_function`` / ``predict_*`` / layer-level kernels on small seeded
inputs and stores inputs, parameters, energy and every gradient.  The shapes
follow the reference's own tests (tests/test_grads_aep.py:18-30,124-135,230-259,
370-410; tests/test_grads_vfe.py:18-39,130-153,371-411) plus BASELINE.json
config 1.  numpy/scipy versions are recorded in each file's ``meta``.
"""
import io
import json
import os
import sys
import contextlib

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import make_ref  # noqa: E402

aep, vfe, lik, kern, utils = make_ref.import_ref()


def perturb(params, rng, scale=0.05):
    out = {}
    for k, v in params.items():
        v = np.array(v, dtype=np.float64)
        out[k] = v + scale * rng.standard_normal(v.shape)
    return out


def save(name, meta, inputs, params, energy, grads, extra=None):
    d = {'meta': json.dumps(dict(meta, numpy=np.__version__, scipy=scipy.__version__,
                                 floor=dict(FLOORS)))}
    for k, v in inputs.items():
        d['in__' + k] = np.asarray(v)
    for k, v in params.items():
        d['p__' + k] = np.asarray(v)
    for k, v in grads.items():
        d['g__' + k] = np.asarray(v)
    d['energy'] = np.asarray(energy, dtype=np.float64).reshape(-1)[:1]
    for k, v in (extra or {}).items():
        d['x__' + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)
    print('%-28s energy=%.10g  keys=%s' % (name, float(d['energy'][0]), sorted(grads)))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return fn(*a, **k)


FLOORS = {}


def run(model, params, mb, alpha, seed=None, prop_mode=None):
    """Evaluate the reference; also estimate its own conditioning floor: how much
    each output moves when every parameter is perturbed by ~1e-15 relative (three
    draws).  Parity cannot be tighter than that, whatever the implementation."""
    import copy
    p = copy.deepcopy(params)
    if seed is not None:
        np.random.seed(seed)
    kw = {} if prop_mode is None else {'prop_mode': prop_mode}
    e, g = model.objective_function(p, mb, alpha=alpha, **kw)
    e = np.array(e, dtype=np.float64).copy()
    g = {k: np.array(v, dtype=np.float64).copy() for k, v in g.items()}
    rng = np.random.RandomState(999)
    floor = {k: 0.0 for k in g}
    floor['energy'] = 0.0
    for _ in range(3):
        q = {k: np.array(v, dtype=np.float64) * (1.0 + 1e-15 * rng.standard_normal(np.shape(v)))
             for k, v in params.items()}
        if seed is not None:
            np.random.seed(seed)
        e2, g2 = model.objective_function(q, mb, alpha=alpha, **kw)
        floor['energy'] = max(floor['energy'], float(np.max(np.abs(e2 - e)) / np.max(np.abs(e))))
        for k in g:
            sc = max(np.max(np.abs(g[k])), 1e-300)
            floor[k] = max(floor[k], float(np.max(np.abs(np.asarray(g2[k]) - g[k])) / sc))
    FLOORS.clear()
    FLOORS.update(floor)
    return e, g


def case_sgpr(name, cls, N, M, D, Do, alpha, nat, mb=None, seed=0, xy=None, lk='Gaussian'):
    rng = np.random.RandomState(seed)
    if xy is None:
        x = rng.standard_normal((N, D))
        y = rng.standard_normal((N, Do))
    else:
        x, y = xy
    if lk == 'Probit':       # binary labels, tests/test_grads_aep.py:138-146
        y = 2.0 * (y > 0) - 1.0
    np.random.seed(seed)
    model = cls(x, y, M, lik=lk, nat_param=nat)
    params = perturb(quiet(model.init_hypers, y), rng)
    if lk == 'Gaussian':
        params['sn'] = np.array(np.log(0.3) + 0.05 * rng.standard_normal())
    mbs = N if mb is None else mb
    e, g = run(model, params, mbs, alpha, seed=123)
    extra = {}
    if mb is None and lk == 'Gaussian':
        xs = rng.standard_normal((7, D))
        model.update_hypers(params)
        model.updated = False
        mf, vf = model.predict_f(xs)
        my, vy = model.predict_y(xs)
        np.random.seed(555)
        fs = model.sample_f(xs, 2)                      # base_models.py:1000-1018, 428-452
        extra = {'xs': xs, 'mf': mf, 'vf': vf, 'my': my, 'vy': vy, 'fs': fs}
    save(name, dict(model=cls.__module__.split('.')[-1] + '.SGPR', N=N, M=M, D=D, Do=Do,
                    alpha=alpha, nat_param=nat, mb_size=mbs, rng_seed=123, lik=lk),
         {'x': x, 'y': y}, params, e, g, extra)


def case_sdgpr(name, N, M, D, hidden, Do, alpha, mb=None, seed=1, lk='Gaussian'):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, D))
    y = rng.standard_normal((N, Do))
    if lk == 'Probit':
        y = 2.0 * (y > 0) - 1.0
    np.random.seed(seed)
    model = aep.SDGPR(x, y, M, hidden, lik=lk)
    params = perturb(quiet(model.init_hypers, y), rng)
    if lk == 'Gaussian':
        params['sn'] = np.array(np.log(0.3))
    mbs = N if mb is None else mb
    e, g = run(model, params, mbs, alpha, seed=123)
    if lk == 'Probit':
        save(name, dict(model='aep_models.SDGPR', N=N, M=M, D=D, hidden=hidden, Do=Do, alpha=alpha,
                        mb_size=mbs, rng_seed=123, lik=lk), {'x': x, 'y': y}, params, e, g)
        return
    xs = rng.standard_normal((6, D))
    model.update_hypers(params)
    model.updated = False
    mf, vf = model.predict_f(xs)
    my, vy = model.predict_y(xs)
    np.random.seed(556)
    smp, mmc, vmc = model.predict_f(xs, prop_mode='MC', no_samples=3)     # base_models.py:1160-1184
    np.random.seed(557)
    fs = model.sample_f(xs, 2)                                            # base_models.py:1239-1262
    save(name, dict(model='aep_models.SDGPR', N=N, M=M, D=D, hidden=hidden, Do=Do, alpha=alpha,
                    mb_size=mbs, rng_seed=123),
         {'x': x, 'y': y}, params, e, g, {'xs': xs, 'mf': mf, 'vf': vf, 'my': my, 'vy': vy,
                                          'mc_samples': smp, 'mc_mf': mmc, 'mc_vf': vmc, 'fs': fs})


def lvm_params(model, y, rng):
    """Hand-built parameters (skips the nested L-BFGS fit of base_models.py:867-871
    but follows the same recipe: base_models.py:853-859, 563-588)."""
    N, Q = y.shape[0], model.Din
    post_m = rng.standard_normal((N, Q))
    post_v = 0.1 * np.ones((N, Q))
    p = quiet(model.sgp_layer.init_hypers, post_m)
    p = perturb(p, rng)
    if getattr(model, '_gold_lik', 'Gaussian') == 'Gaussian':
        p['sn'] = np.array(np.log(0.3))
    if model.nat_param:
        post_2 = 1.0 / post_v
        p['x1'] = post_2 * post_m
        p['x2'] = np.log(post_2 - 1) / 2 + 0.05 * rng.standard_normal((N, Q))
    else:
        p['x1'] = post_m
        p['x2'] = np.log(post_v) / 2 + 0.05 * rng.standard_normal((N, Q))
    return p


def case_sgplvm(name, cls, N, M, Q, Do, alpha, nat=True, mb=None, seed=2, lk='Gaussian', prop_mode=None):
    rng = np.random.RandomState(seed)
    y = rng.standard_normal((N, Do))
    if lk == 'Probit':       # tests/test_grads_aep.py:33-40
        y = 2.0 * (y > 0) - 1.0
    np.random.seed(seed)
    model = cls(y, Q, M, lik=lk, nat_param=nat)
    model._gold_lik = lk
    params = lvm_params(model, y, rng)
    mbs = N if mb is None else mb
    e, g = run(model, params, mbs, alpha, seed=123, prop_mode=prop_mode)
    meta = dict(model=cls.__module__.split('.')[-1] + '.SGPLVM', N=N, M=M, Q=Q, Do=Do,
                alpha=alpha, nat_param=nat, mb_size=mbs, rng_seed=123, lik=lk)
    if prop_mode is not None:
        meta['prop_mode'] = prop_mode
    save(name, meta, {'y': y}, params, e, g)


def ssm_params(model, y, rng, gp_emi):
    N, Q = y.shape[0], model.Din
    post_m = y[:, :Q] + 0.1 * rng.standard_normal((N, Q)) if y.shape[1] >= Q else rng.standard_normal((N, Q))
    post_v = 0.1 * np.ones_like(post_m)
    p = {'sn': np.array([np.log(0.2)])}
    if model.nat_param:
        post_2 = 1.0 / post_v
        p['x_factor_1'] = post_2 * post_m / 3
        p['x_factor_2'] = np.log(post_2 / 3) / 2 + 0.05 * rng.standard_normal((N, Q))
    else:
        p['x_factor_1'] = post_m.copy()
        p['x_factor_2'] = np.log(post_v) / 2 + 0.05 * rng.standard_normal((N, Q))
    xin = post_m[:N - 1]
    if model.Dcon_dyn > 0:
        xin = np.hstack((xin, model.x_control[:N - 1]))
    p.update(perturb(quiet(model.dyn_layer.init_hypers, xin, key_suffix='_dynamic'), rng))
    if gp_emi:
        xin = post_m
        if model.Dcon_emi > 0:
            xin = np.hstack((xin, model.x_control))
        p.update(perturb(quiet(model.emi_layer.init_hypers, xin, key_suffix='_emission'), rng))
        p['sn_emission'] = np.array(np.log(0.25))
    else:
        Dout, Din = model.Dout, Q + model.Dcon_emi
        p['C_emission'] = np.eye(Dout, Din) + 0.1 * rng.standard_normal((Dout, Din))
        p['R_emission'] = np.log(0.2) * np.ones(Dout) + 0.05 * rng.standard_normal(Dout)
    return p


def case_sgpssm(name, cls, N, M, Q, Do, alpha, gp_emi=False, control=0, nat=True, mb=None, seed=3,
                prop_mode=None, predict=False):
    rng = np.random.RandomState(seed)
    y = np.cumsum(0.3 * rng.standard_normal((N, Do)), axis=0)
    xc = rng.standard_normal((N, control)) if control else None
    np.random.seed(seed)
    if cls is aep.SGPSSM:
        model = cls(y, Q, M, lik='Gaussian', x_control=xc, gp_emi=gp_emi)
    else:
        model = cls(y, Q, M, lik='Gaussian', x_control=xc, gp_emi=gp_emi, nat_param=nat)
    params = ssm_params(model, y, rng, gp_emi)
    mbs = N if mb is None else mb
    e, g = run(model, params, mbs, alpha, seed=123, prop_mode=prop_mode)
    inputs = {'y': y}
    if control:
        inputs['x_control'] = xc
    meta = dict(model=cls.__module__.split('.')[-1] + '.SGPSSM', N=N, M=M, Q=Q, Do=Do,
                alpha=alpha, gp_emi=gp_emi, control=control, nat_param=nat, mb_size=mbs,
                rng_seed=123)
    if prop_mode is not None:
        meta['prop_mode'] = prop_mode
    extra = {}
    if predict:
        # prediction API around the path (base_models.py:1453-1595): T-step roll-outs from the last
        # state (moment matching / particles), one-step predict_y, posterior over the observations
        model.update_hypers(params)
        Tf = 4
        xcf = rng.standard_normal((Tf, control)) if control else None
        pf = model.predict_forward_mm(Tf, xcf)
        xin = rng.standard_normal((5, Q + control))
        py = model.predict_y(xin)
        gy = model.get_posterior_y()
        np.random.seed(321)
        pmc = model.predict_forward_mc(3, xcf, 4) if not (control and not gp_emi) else None
        extra = {'pf_mx': pf[0], 'pf_vx': pf[1], 'pf_my': pf[2], 'pf_vyn': pf[3], 'pf_vy': pf[4],
                 'py_in': xin, 'py_my': py[0], 'py_vy': py[1], 'gy_my': gy[0], 'gy_vf': gy[1], 'gy_vyn': gy[2]}
        if xcf is not None:
            extra['pf_xc'] = xcf
        if pmc is not None:
            extra.update({'mc_x': pmc[0], 'mc_my': pmc[1], 'mc_vy': pmc[2]})
    save(name, meta, inputs, params, e, g, extra)


def case_kernels(seed=4):
    rng = np.random.RandomState(seed)
    N, M, Q = 9, 6, 3
    mx = rng.standard_normal((N, Q))
    vx = rng.rand(N, Q) * 0.5 + 0.01
    z = rng.standard_normal((M, Q))
    ls = 0.3 * rng.standard_normal(Q)
    sf = np.array([0.2])
    kfu = kern.compute_kernel(2 * ls, 2 * sf, mx, z)
    psi1, psi2 = kern.compute_psi_weave(2 * ls, 2 * sf, mx, vx, z)
    psi1n = kern.psi1computations(np.exp(2 * sf), np.exp(ls), z, mx, vx)
    psi2n = kern.psi2computations(np.exp(2 * sf), np.exp(ls), z, mx, vx)
    dpsi1 = rng.standard_normal((N, M))
    dpsi2 = rng.standard_normal((N, M, M))
    der = kern.compute_psi_derivatives(dpsi1, psi1, dpsi2, psi2, np.exp(ls), np.exp(2 * sf), mx, vx, z)
    dk = kern.compute_kfu_derivatives(dpsi1, kfu, np.exp(ls), np.exp(2 * sf), mx, z, grad_x=True)
    Mm = rng.standard_normal((M, M))
    Kzz = kern.compute_kernel(2 * ls, 2 * sf, z, z)
    tr = kern.d_trace_MKzz_dhypers(2 * ls, 2 * sf, z, Mm, Kzz)
    d = dict(mx=mx, vx=vx, z=z, ls=ls, sf=sf, kfu=kfu, psi1=psi1, psi2=psi2, psi1_numpy=psi1n,
             psi2_numpy=psi2n, dpsi1=dpsi1, dpsi2=dpsi2, Mm=Mm, Kzz=Kzz,
             psider_var=der[0], psider_l=der[1], psider_z=der[2], psider_mu=der[3], psider_S=der[4],
             kfuder_var=dk[0], kfuder_l=dk[1], kfuder_z=dk[2], kfuder_x=dk[3],
             tr_sf=tr[0], tr_ls=tr[1], tr_z=tr[2],
             meta=json.dumps(dict(numpy=np.__version__, scipy=scipy.__version__)))
    np.savez_compressed(os.path.join(HERE, 'kernels.npz'), **d)
    print('kernels.npz written')


def case_emis(seed=5):
    """tests/test_grads_emis.py shapes: Gauss_Emis tilted + log-lik-exp."""
    rng = np.random.RandomState(seed)
    N, Do, Q = 8, 3, 2
    y = rng.standard_normal((N, Do))
    em = lik.Gauss_Emis(y, Do, Q)
    p = {'C': rng.standard_normal((Do, Q)), 'R': 0.3 * rng.standard_normal(Do)}
    em.update_hypers(p)
    mx = rng.standard_normal((N, Q))
    vx = rng.rand(N, Q) + 0.05
    lz, gi, gh = em.compute_emission_tilted(mx, vx, 0.7, -1.3)
    le, gi2, gh2 = em.compute_emission_log_lik_exp(mx, vx, -1.3)
    np.savez_compressed(os.path.join(HERE, 'gauss_emis.npz'), y=y, C=p['C'], R=p['R'], mx=mx, vx=vx,
                        alpha=0.7, scale=-1.3, t_logZ=lz, t_dmx=gi['mx'], t_dvx=gi['vx'],
                        t_dC=gh['C'], t_dR=gh['R'], e_logZ=le, e_dmx=gi2['mx'], e_dvx=gi2['vx'],
                        e_dC=gh2['C'], e_dR=gh2['R'],
                        meta=json.dumps(dict(numpy=np.__version__, scipy=scipy.__version__)))
    print('gauss_emis.npz written')


def probit_cases():
    # tests/test_grads_aep.py:33-40,138-146,246-256; tests/test_grads_vfe.py probit twins
    case_sgpr('aep_sgpr_probit', aep.SGPR, 20, 10, 2, 3, 0.5, True, seed=50, lk='Probit')
    case_sgpr('aep_sgpr_probit_alpha_one', aep.SGPR, 20, 10, 2, 3, 1.0, True, seed=51, lk='Probit')
    case_sgpr('vfe_sgpr_probit', vfe.SGPR, 20, 10, 2, 3, 1.0, True, seed=52, lk='Probit')
    case_sdgpr('aep_sdgpr_probit', 10, 5, 2, [3, 2], 2, 0.5, seed=53, lk='Probit')
    case_sgplvm('aep_sgplvm_probit', aep.SGPLVM, 10, 5, 3, 2, 0.5, seed=54, lk='Probit')
    case_sgplvm('vfe_sgplvm_probit', vfe.SGPLVM, 10, 5, 3, 2, 1.0, seed=55, lk='Probit')


def case_sdgprh(name, N, M, D, hidden, Do, alpha, seed=60, lk='Gaussian', init_recipe=True):
    """tests/test_grads_aep.py:340-367 (SDGPR_H: deep GP with per-row hidden-variable factors)."""
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((N, D))
    y = rng.standard_normal((N, Do))
    if lk == 'Probit':
        y = 2.0 * (y > 0) - 1.0
    np.random.seed(seed)
    model = aep.SDGPR_H(x, y, M, hidden, lik=lk)
    params = perturb(quiet(model.init_hypers, y), rng)
    if lk == 'Gaussian':
        params['sn'] = np.array(np.log(0.3))
    if not init_recipe:      # moderate factor precisions / transition noise instead of 1e-4 / 1e-3
        size = [D] + list(hidden) + [Do]
        params['sn_hidden'] = np.log(0.2) + 0.05 * rng.standard_normal(len(hidden))
        for i in range(len(hidden)):
            params['h_factor_1_%d' % i] = 0.5 * rng.standard_normal((N, size[i + 1]))
            params['h_factor_2_%d' % i] = np.log(1.5) / 2 + 0.1 * rng.standard_normal((N, size[i + 1]))
    e, g = run(model, params, N, alpha, seed=123)
    extra = {}
    if lk == 'Gaussian':
        xs = rng.standard_normal((6, D))
        model.update_hypers(params)
        model.updated = False
        mf, vf = model.predict_f(xs)
        my, vy = model.predict_y(xs)
        extra = {'xs': xs, 'mf': mf, 'vf': vf, 'my': my, 'vy': vy}
    save(name, dict(model='aep_models.SDGPR_H', N=N, M=M, D=D, hidden=hidden, Do=Do, alpha=alpha,
                    mb_size=N, rng_seed=123, lik=lk), {'x': x, 'y': y}, params, e, g, extra)


def case_input_grad(seed=70):
    """predict_f_with_input_grad / predict_y_with_input_grad (base_models.py:1186-1237,1277-1289) and
    the layer-level backprop_predictive_grads_reg (391-426) of single-layer deep GPs (the only depth
    the reference defines them for), scalar and 2-D outputs."""
    out = {}
    meta = dict(model='aep_models.SDGPR', cases=[])
    for tag, N, M, D, Do in (('a', 12, 5, 3, 1), ('b', 9, 4, 2, 2)):
        rng = np.random.RandomState(seed + Do)
        x = rng.standard_normal((N, D))
        y = rng.standard_normal((N, Do))
        np.random.seed(seed)
        model = aep.SDGPR(x, y, M, [], lik='Gaussian')
        params = perturb(quiet(model.init_hypers, y), rng)
        params['sn'] = np.array(np.log(0.3))
        xs = rng.standard_normal((7, D))
        model.update_hypers(params)
        model.updated = False
        mf, vf, dm_dx, dv_dx = model.predict_f_with_input_grad(xs)
        my, vy, _, _ = model.predict_y_with_input_grad(xs)
        layer = model.sgp_layers[0]
        m0, v0, kfu = layer.forward_prop_thru_post(xs, return_info=True)
        w_m = rng.standard_normal((1, Do))
        w_v = rng.standard_normal((1, Do))
        l_dm, l_dv = layer.backprop_predictive_grads_reg(m0, v0, w_m, w_v, np.zeros((1, 1)), np.ones((1, 1)), kfu, xs)
        meta['cases'].append(dict(tag=tag, N=N, M=M, D=D, Do=Do))
        for k, v in params.items():
            out['%s_p_%s' % (tag, k)] = np.asarray(v)
        for k, v in dict(x=x, y=y, xs=xs, mf=mf, vf=vf, dm_dx=dm_dx, dv_dx=dv_dx, my=my, vy=vy,
                         w_m=w_m, w_v=w_v, l_dm=l_dm, l_dv=l_dv).items():
            out['%s_%s' % (tag, k)] = np.asarray(v)
    meta.update(numpy=np.__version__, scipy=scipy.__version__)
    np.savez_compressed(os.path.join(HERE, 'input_grad.npz'), meta=json.dumps(meta), **out)
    print('wrote input_grad')


def case_layer_iface(seed=80):
    """Layer-level interface of SURVEY 8b on small shapes: compute_cavity / forward_prop_thru_cav /
    backprop_grads_reg / backprop_grads_lvm_mm / compute_phi (+ prior, posterior, cavity) /
    forward_prop_thru_post of aep.SGP_Layer (aep_models.py:62-304,413-546), compute_KL and the two
    backprops of vfe.SGP_Layer (vfe_models.py:309-401,479-548), natural and non-natural parameters."""
    out = {}
    meta = dict(cases=[])
    for tag, mod, nat, alpha in (('aep_nat', 'aep', True, 0.6), ('aep_non', 'aep', False, 0.3),
                                 ('vfe_nat', 'vfe', True, 1.0), ('vfe_non', 'vfe', False, 1.0)):
        N, M, D, Do, n = 30, 6, 3, 2, 8
        rng = np.random.RandomState(seed + len(meta['cases']))
        np.random.seed(seed)
        cls = aep.SGP_Layer if mod == 'aep' else vfe.SGP_Layer
        layer = cls(N, D, Do, M, nat)
        xtr = rng.standard_normal((N, D))
        params = perturb(quiet(layer.init_hypers, xtr), rng)
        layer.update_hypers(params)
        layer.compute_kuu()
        layer.update_posterior()
        r = {}
        x = rng.standard_normal((n, D))
        mx = rng.standard_normal((n, D))
        vx = 0.1 + rng.rand(n, D)
        dm, dv = rng.standard_normal((n, Do)), rng.standard_normal((n, Do))
        dm2, dv2 = rng.standard_normal((n, Do)), rng.standard_normal((n, Do))
        if mod == 'aep':
            layer.compute_cavity(alpha)
            r['phi'] = np.array([layer.compute_phi(alpha), layer.compute_phi_prior(),
                                 layer.compute_phi_posterior(), layer.compute_phi_cavity()])
            m, v, kfu = layer.forward_prop_thru_cav(x)
            g = layer.backprop_grads_reg(m, v, dm, dv, kfu, x, alpha)
            ms, vs, psi1, psi2 = layer.forward_prop_thru_cav(mx, vx, mode='MM')
            gs, gx = layer.backprop_grads_lvm_mm(ms, vs, dm2, dv2, psi1, psi2, mx, vx, alpha)
        else:
            r['phi'] = np.array([layer.compute_KL()])
            m, v, kfu = layer.forward_prop_thru_post(x, return_info=True)
            g = layer.backprop_grads_reg(m, v, dm, dv, kfu, x)
            ms, vs, psi1, psi2 = layer.forward_prop_thru_post(mx, vx, mode='MM', return_info=True)
            gs, gx = layer.backprop_grads_lvm_mm(ms, vs, dm2, dv2, psi1, psi2, mx, vx)
        pm, pv = layer.forward_prop_thru_post(x)
        pms, pvs = layer.forward_prop_thru_post(mx, vx, mode='MM')
        # Monte-Carlo propagation at layer level (aep_models.py:160-180,307-410; base_models.py:309-332,
        # 373-388; vfe_models.py:405-476); the reference's AEP chain rule is natural-parameter only
        if not (mod == 'aep' and not nat):
            np.random.seed(321)
            if mod == 'aep':
                res, res_s = layer.forward_prop_thru_cav(mx, vx, mode='MC')
            else:
                res, res_s = layer.forward_prop_thru_post(mx, vx, mode='MC', return_info=True)
            K = res[0].shape[0]
            dm3, dv3 = rng.standard_normal((K, n, Do)), rng.standard_normal((K, n, Do))
            if mod == 'aep':
                gmc, dxs = layer.backprop_grads_lvm_mc(res_s[0], res_s[1], dm3, dv3, res_s[2], res_s[3], alpha)
            else:
                gmc, dxs = layer.backprop_grads_lvm_mc(res_s[0], res_s[1], dm3, dv3, res_s[2], res_s[3])
            gin = layer.backprop_grads_reparam(dxs, mx, vx, res[4])
            r.update(dm3=dm3, dv3=dv3, mc_m=res[0], mc_v=res[1], mc_kfu=res[2], mc_x=res[3], mc_eps=res[4],
                     mc_dxs=dxs, mc_gx_mx=gin['mx'], mc_gx_vx=gin['vx'])
            for k, a in gmc.items():
                r['mcg_' + k] = a
            np.random.seed(322)
            pmc = layer.forward_prop_thru_post(mx, vx, mode='MC')
            r.update(mc_pm=pmc[0], mc_pv=pmc[1])
        r.update(xtr=xtr, x=x, mx=mx, vx=vx, dm=dm, dv=dv, dm2=dm2, dv2=dv2, m=m, v=v, kfu=kfu, ms=ms, vs=vs,
                 psi1=psi1, psi2=psi2, pm=pm, pv=pv, pms=pms, pvs=pvs, gx_mx=gx['mx'], gx_vx=gx['vx'])
        for k, a in params.items():
            r['p_' + k] = a
        for k, a in g.items():
            r['g_' + k] = a
        for k, a in gs.items():
            r['gs_' + k] = a
        for k, a in r.items():
            out[tag + '__' + k] = np.asarray(a)
        meta['cases'].append(dict(tag=tag, mod=mod, nat=nat, alpha=alpha, N=N, M=M, D=D, Do=Do))
    meta.update(numpy=np.__version__, scipy=scipy.__version__)
    np.savez_compressed(os.path.join(HERE, 'layer_iface.npz'), meta=json.dumps(meta), **out)
    print('wrote layer_iface')


def case_lik_iface(seed=90):
    """Public likelihood-layer interface (lik_layers.py:104-236 Gauss_Layer, 303-457 Probit_Layer):
    compute_log_Z / backprop_grads / compute_log_lik_exp / backprop_grads_log_lik_exp on 2-D inputs and
    on the 3-D inputs of Monte-Carlo propagation, alpha = 0.4 and 1, plus Gauss compute_dm2."""
    rng = np.random.RandomState(seed)
    n, D, K = 7, 2, 4
    out, cases = {}, []
    for name in ('Gauss', 'Probit'):
        for alpha in (0.4, 1.0):
            for shape in ((n, D), (K, n, D)):
                tag = '%s_a%g_%dd' % (name, alpha, len(shape))
                L = getattr(lik, name + '_Layer')(n, D)
                if name == 'Gauss':
                    L.update_hypers({'sn': np.array(np.log(0.3))})
                    y = rng.standard_normal((n, D))
                else:
                    y = 2.0 * (rng.standard_normal((n, D)) > 0) - 1
                m, v = rng.standard_normal(shape), 0.2 + rng.rand(*shape)
                v1 = v.copy()
                r = L.compute_log_Z(m, v1, y, alpha)
                g = L.backprop_grads(m, v1, r[1], r[2], alpha, 0.7)
                e = L.compute_log_lik_exp(m, v, y)
                ge = L.backprop_grads_log_lik_exp(m, v, e[1], e[2], y, 0.7)
                d = dict(y=y, m=m, v=v, vout=v1, logZ=np.array([r[0]]), dm=r[1], dv=r[2], ll=np.array([e[0]]),
                         edm=e[1], edv=e[2])
                if name == 'Gauss':
                    d.update(g_sn=np.array([g['sn']]), ge_sn=np.array([ge['sn']]))
                    if len(shape) == 2:
                        d['dm2'] = L.compute_log_Z(m, v.copy(), y, alpha, compute_dm2=True)[3]
                else:
                    assert g == {} and ge == {}
                for k, a in d.items():
                    out[tag + '__' + k] = np.asarray(a)
                cases.append(dict(tag=tag, lik=name, alpha=alpha, n=n, D=D))
    meta = dict(cases=cases, sn=float(np.log(0.3)), numpy=np.__version__, scipy=scipy.__version__)
    np.savez_compressed(os.path.join(HERE, 'lik_iface.npz'), meta=json.dumps(meta), **out)
    print('wrote lik_iface')


def sdgprh_cases():
    case_sdgprh('aep_sdgprh', 10, 5, 2, [3, 2], 3, 0.5)
    case_sdgprh('aep_sdgprh_moderate', 12, 6, 3, [2, 2], 2, 0.7, seed=61, init_recipe=False)
    case_sdgprh('aep_sdgprh_alpha_one', 10, 5, 2, [3], 1, 1.0, seed=62, init_recipe=False)
    case_sdgprh('aep_sdgprh_probit', 8, 4, 2, [3, 2], 3, 0.3, seed=63, lk='Probit', init_recipe=False)


def mc_cases():
    """Monte-Carlo propagation (config.PROP_MC): eps comes from the global numpy RNG, which the
    harness re-seeds before every evaluation (as tests/test_utils.py:70-72 does for stochastic runs)."""
    case_sgplvm('aep_sgplvm_mc', aep.SGPLVM, 10, 5, 3, 2, 0.5, seed=70, prop_mode='MC')
    case_sgplvm('aep_sgplvm_mc_minibatch', aep.SGPLVM, 12, 6, 2, 3, 0.8, mb=5, seed=71, prop_mode='MC')
    case_sgplvm('vfe_sgplvm_mc', vfe.SGPLVM, 10, 5, 3, 2, 1.0, seed=72, prop_mode='MC')
    case_sgpssm('aep_sgpssm_lin_mc', aep.SGPSSM, 20, 4, 2, 2, 0.5, seed=73, prop_mode='MC')
    case_sgpssm('aep_sgpssm_gp_mc', aep.SGPSSM, 10, 4, 2, 3, 0.5, gp_emi=True, seed=74, prop_mode='MC')
    case_sgplvm('aep_sgplvm_probit_mc', aep.SGPLVM, 10, 5, 3, 2, 0.5, seed=78, lk='Probit', prop_mode='MC')
    case_sgplvm('aep_sgplvm_probit_mc_alpha_one', aep.SGPLVM, 10, 5, 2, 2, 1.0, seed=79, lk='Probit', prop_mode='MC')
    case_sgplvm('vfe_sgplvm_probit_mc', vfe.SGPLVM, 10, 5, 3, 2, 1.0, seed=80, lk='Probit', prop_mode='MC')
    case_sgpssm('vfe_sgpssm_lin_mc', vfe.SGPSSM, 20, 4, 2, 2, 1.0, seed=76, prop_mode='MC')
    case_sgpssm('vfe_sgpssm_gp_mc', vfe.SGPSSM, 10, 4, 2, 3, 1.0, gp_emi=True, seed=77, prop_mode='MC')
    case_sgpssm('aep_sgpssm_control_mc', aep.SGPSSM, 12, 4, 2, 2, 0.7, control=1, mb=7, seed=75, prop_mode='MC')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'probit':   # only the files added with the probit layer
        probit_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'sample':
        case_sgpr('aep_sgpr', aep.SGPR, 20, 10, 2, 3, 0.5, True)
        case_sdgpr('aep_sdgpr', 10, 5, 2, [3, 2], 2, 1.0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'ssm_predict':
        case_sgpssm('aep_sgpssm_lin', aep.SGPSSM, 20, 4, 2, 2, 0.5, predict=True)
        case_sgpssm('aep_sgpssm_gp', aep.SGPSSM, 10, 4, 2, 3, 0.5, gp_emi=True, seed=42, predict=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'input_grad':
        case_input_grad()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'layer_iface':
        case_layer_iface()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'lik_iface':
        case_lik_iface()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'mc':
        mc_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'sdgprh':   # only the files added with SDGPR_H
        sdgprh_cases()
        sys.exit(0)
    # tests/test_grads_aep.py:124-135 shape (alpha 0.5) + its alpha=1e-4 + non-natural params
    case_sgpr('aep_sgpr', aep.SGPR, 20, 10, 2, 3, 0.5, True)
    case_sgpr('aep_sgpr_alpha_small', aep.SGPR, 20, 10, 2, 3, 1e-4, True, seed=10)
    case_sgpr('aep_sgpr_alpha_one', aep.SGPR, 20, 10, 2, 3, 1.0, True, seed=11)
    case_sgpr('aep_sgpr_nonnat', aep.SGPR, 20, 10, 2, 3, 0.5, False, seed=12)
    case_sgpr('aep_sgpr_minibatch', aep.SGPR, 20, 10, 2, 3, 0.5, True, mb=7, seed=13)
    case_sgpr('vfe_sgpr', vfe.SGPR, 20, 10, 2, 3, 0.5, True, seed=14)
    case_sgpr('vfe_sgpr_nonnat', vfe.SGPR, 20, 10, 2, 3, 0.5, False, seed=15)
    # BASELINE.json config 1: examples/gpr_aep_examples.py:12-18 data, M=50, alpha=0.5
    rs = np.random.RandomState(42)
    X = rs.rand(200, 1)
    Y = np.sin(12 * X) + 0.5 * np.cos(25 * X) + rs.randn(200, 1) * 0.2
    case_sgpr('aep_sgpr_cfg1', aep.SGPR, 200, 50, 1, 1, 0.5, True, seed=16, xy=(X, Y))
    # tests/test_grads_aep.py:230-259
    case_sdgpr('aep_sdgpr', 10, 5, 2, [3, 2], 2, 1.0)
    case_sdgpr('aep_sdgpr_alpha_half', 12, 6, 3, [2, 2], 1, 0.5, seed=20)
    case_sdgpr('aep_sdgpr_minibatch', 12, 6, 3, [2], 1, 0.5, mb=5, seed=21)
    # tests/test_grads_aep.py:18-30; tests/test_grads_vfe.py:18-39
    case_sgplvm('aep_sgplvm', aep.SGPLVM, 10, 5, 3, 2, 0.5)
    case_sgplvm('aep_sgplvm_minibatch', aep.SGPLVM, 12, 5, 2, 3, 0.7, mb=5, seed=30)
    case_sgplvm('vfe_sgplvm', vfe.SGPLVM, 10, 5, 3, 2, 1.0, seed=31)
    case_sgplvm('vfe_sgplvm_nonnat', vfe.SGPLVM, 10, 5, 3, 2, 1.0, nat=False, seed=32)
    # tests/test_grads_aep.py:370-410; tests/test_grads_vfe.py:371-411
    case_sgpssm('aep_sgpssm_lin', aep.SGPSSM, 20, 4, 2, 2, 0.5, predict=True)
    case_sgpssm('aep_sgpssm_lin_1d', aep.SGPSSM, 30, 4, 1, 1, 0.4, seed=40)
    case_sgpssm('aep_sgpssm_lin_window', aep.SGPSSM, 20, 4, 2, 2, 0.5, mb=8, seed=41)
    case_sgpssm('aep_sgpssm_gp', aep.SGPSSM, 10, 4, 2, 3, 0.5, gp_emi=True, seed=42, predict=True)
    case_sgpssm('aep_sgpssm_control', aep.SGPSSM, 12, 4, 2, 2, 0.5, control=1, seed=43)
    case_sgpssm('vfe_sgpssm_lin', vfe.SGPSSM, 20, 4, 2, 2, 1.0, seed=44)
    case_sgpssm('vfe_sgpssm_gp', vfe.SGPSSM, 10, 4, 2, 3, 1.0, gp_emi=True, seed=45)
    case_sgpssm('vfe_sgpssm_nonnat', vfe.SGPSSM, 12, 4, 2, 2, 1.0, nat=False, seed=46)
    probit_cases()
    case_kernels()
    case_emis()
    probit_cases()
    sdgprh_cases()
    mc_cases()
    case_input_grad()
    case_layer_iface()
    case_lik_iface()
