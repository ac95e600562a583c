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
