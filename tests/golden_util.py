"""Helpers shared by the parity tests: load tests/golden/*.npz (generated from the
reference by tests/golden/gen_golden.py) and compare gradient dicts."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    f = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    out = {'meta': json.loads(str(f['meta'])), 'in': {}, 'p': {}, 'g': {}, 'x': {}}
    for k in f.files:
        for pre in ('in', 'p', 'g', 'x'):
            if k.startswith(pre + '__'):
                out[pre][k[len(pre) + 2:]] = np.array(f[k])
    if 'energy' in f.files:
        out['energy'] = float(f['energy'][0])
    return out


def model_cases(prefix=None, bench=False):
    """Golden model cases.  `bench_*` files are reduced-n runs at the benchmark's own shapes
    (M = 128 ... 512; tests/golden/gen_golden_bench.py): too slow for the CPU fiber emulator and
    the numpy oracle, so they are listed only when `bench` is set (GPU tests)."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))
    names = [n for n in names if n not in ('kernels', 'gauss_emis', 'input_grad', 'layer_iface', 'lik_iface')]
    names = [n for n in names if n.startswith('bench_') == bool(bench)]
    if prefix:
        names = [n for n in names if n.startswith(prefix)]
    return names


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): relative to the key's scale (a gradient
    entry that is ~0 by cancellation cannot be matched to 1e-6 of itself)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape or a.size == b.size, (a.shape, b.shape)
    a = a.reshape(-1)
    b = b.reshape(-1)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if b.size else 0.0


def assert_close(energy, grads, gold, tol, what=''):
    """Energy and every gradient key within ``tol`` relative -- except where the
    reference's OWN output is less certain than that: ``meta['floor'][key]`` is how
    far the reference's result moves under a 1e-15 relative perturbation of its
    inputs (tests/golden/gen_golden.py); the bound used is max(tol, 10*floor)."""
    floor = gold['meta'].get('floor', {})
    e = float(np.ravel(energy)[0])
    t = max(tol, 10 * floor.get('energy', 0.0))
    assert abs(e - gold['energy']) <= t * max(abs(gold['energy']), 1e-12), \
        '%s energy %r vs %r' % (what, e, gold['energy'])
    assert set(grads.keys()) == set(gold['g'].keys()), (sorted(grads), sorted(gold['g']))
    for k in gold['g']:
        r = rel_err(grads[k], gold['g'][k])
        t = max(tol, 10 * floor.get(k, 0.0))
        assert r <= t, '%s grad %s rel err %.3e > %.1e' % (what, k, r, t)


def build_oracle_model(gold):
    """Instantiate the oracle model described by a golden file."""
    import geepee_oracle as go
    m = gold['meta']
    i = gold['in']
    kind = m['model']
    lk = m.get('lik', 'Gaussian')
    if kind == 'aep_models.SGPR':
        return go.AepSGPR(i['x'], i['y'], m['M'], m['nat_param'], lik=lk)
    if kind == 'vfe_models.SGPR':
        return go.VfeSGPR(i['x'], i['y'], m['M'], m['nat_param'], lik=lk)
    if kind == 'aep_models.SDGPR':
        return go.AepSDGPR(i['x'], i['y'], m['M'], m['hidden'], lik=lk)
    if kind == 'aep_models.SDGPR_H':
        return go.AepSDGPR_H(i['x'], i['y'], m['M'], m['hidden'], lik=lk)
    if kind == 'aep_models.SGPLVM':
        return go.AepSGPLVM(i['y'], m['Q'], m['M'], lik=lk)
    if kind == 'vfe_models.SGPLVM':
        return go.VfeSGPLVM(i['y'], m['Q'], m['M'], nat_param=m['nat_param'], lik=lk)
    if kind == 'aep_models.SGPSSM':
        return go.AepSGPSSM(i['y'], m['Q'], m['M'], x_control=i.get('x_control'), gp_emi=m['gp_emi'])
    if kind == 'vfe_models.SGPSSM':
        return go.VfeSGPSSM(i['y'], m['Q'], m['M'], x_control=i.get('x_control'),
                            gp_emi=m['gp_emi'], nat_param=m['nat_param'])
    raise ValueError(kind)
