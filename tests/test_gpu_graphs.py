"""The CUDA-graph replay of the replicated tails (geepee_b200/tailgraph.py) must be invisible:
same energy and gradients as the eager launches, for NEW parameter values on every replay
(static input buffers really refreshed), against the golden vectors from the reference."""
import copy

import numpy as np
import pytest
import torch

import golden_util as gu
import model_cases as mc

pytestmark = pytest.mark.gpu

CASES = ['aep_sgpr', 'aep_sgpr_cfg1', 'aep_sgpr_nonnat', 'aep_sdgpr', 'aep_sgplvm', 'aep_sgpssm_lin',
         'aep_sgpssm_gp', 'vfe_sgpr', 'vfe_sgplvm', 'vfe_sgpssm_lin', 'aep_sgpr_probit', 'aep_sdgprh_moderate',
         'aep_sgplvm_mc', 'vfe_sgplvm_mc']


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()
    yield


def _layers(model):
    out = []
    for name in ('sgp_layer', 'dyn_layer', 'emi_layer'):
        layer = getattr(model, name, None)
        if hasattr(layer, '_graphs'):
            out.append(layer)
    return out + list(getattr(model, 'sgp_layers', []))


def _perturbed(p, seed):
    rng = np.random.RandomState(seed)
    return {k: np.array(v, dtype=np.float64) + 1e-2 * rng.standard_normal(np.shape(v)) *
            (1.0 if k.startswith(('eta', 'zu', 'ls', 'sf', 'x', 'C_')) else 0.1) for k, v in p.items()}


def _call(model, gold, p):
    m = gold['meta']
    np.random.seed(m['rng_seed'])
    return model.objective_function(copy.deepcopy(p), m['mb_size'], alpha=m['alpha'],
                                    prop_mode=m.get('prop_mode', 'MM'))


@pytest.mark.parametrize('name', CASES)
def test_replay_matches_eager_and_golden(name, monkeypatch):
    from geepee_b200 import config
    gold = gu.load(name)
    model = mc.build_model(gold)
    warm = config.TAIL_GRAPH_WARMUP
    for _ in range(warm + 1):                    # eager calls, then the capturing call
        e, g = _call(model, gold, gold['p'])
        gu.assert_close(e, g, gold, 1e-6, name + ' (warm-up / capture)')
    layers = _layers(model)
    assert layers and all(L._active_pre is not None and L._active_pre.captured for L in layers), \
        'tails were not captured'
    assert all(any(c.captured for c in L._active_pre.children.values()) for L in layers), \
        'post-tails were not captured'
    # replays with different parameter values against an eager-only twin
    monkeypatch.setattr(config, 'TAIL_GRAPHS', False)
    twin = mc.build_model(gold)
    monkeypatch.setattr(config, 'TAIL_GRAPHS', True)
    for seed in (11, 12):
        p = _perturbed(gold['p'], seed)
        monkeypatch.setattr(config, 'TAIL_GRAPHS', False)
        e0, g0 = _call(twin, gold, p)
        monkeypatch.setattr(config, 'TAIL_GRAPHS', True)
        e1, g1 = _call(model, gold, p)
        assert np.isfinite(e0)
        gu.assert_close(e1, g1, {'energy': e0, 'g': g0, 'meta': {}}, 1e-9, name + ' (replay vs eager)')
    assert all(L._active_pre is not None for L in layers)
    e, g = _call(model, gold, gold['p'])
    gu.assert_close(e, g, gold, 1e-6, name + ' (replay)')


def test_replay_counts_library_launches():
    """gpu_launches (bench.py) stays a count of executed library kernels under replay."""
    from geepee_b200 import ops
    gold = gu.load('aep_sgpr')
    model = mc.build_model(gold)
    counts = []
    for _ in range(5):
        n0 = ops.launch_count()
        _call(model, gold, gold['p'])
        counts.append(ops.launch_count() - n0)
    assert counts[-1] == counts[-2] and counts[-1] >= counts[0] - 2 and counts[-1] > 0, counts


def test_predict_after_replay():
    gold = gu.load('aep_sgpr')
    model = mc.build_model(gold)
    for _ in range(4):
        _call(model, gold, gold['p'])
    x = gold['x']
    mf, vf = model.predict_f(x['xs'])
    assert gu.rel_err(mf, x['mf']) < 1e-7 and gu.rel_err(vf, x['vf']) < 1e-7


def test_alpha_switching_keeps_graphs_apart(monkeypatch):
    """One captured pre-/post-tail pair per alpha: switching back and forth must replay the right
    one (the cavity and the log-partition scales depend on alpha)."""
    from geepee_b200 import config
    gold = gu.load('aep_sgpr')
    model = mc.build_model(gold)
    monkeypatch.setattr(config, 'TAIL_GRAPHS', False)
    twin = mc.build_model(gold)
    ref = {}
    for alpha in (0.5, 1.0, 0.25):
        ref[alpha] = twin.objective_function(copy.deepcopy(gold['p']), gold['meta']['mb_size'], alpha=alpha)
    monkeypatch.setattr(config, 'TAIL_GRAPHS', True)
    for alpha in (0.5, 0.5, 0.5, 0.5, 1.0, 1.0, 1.0, 1.0, 0.5, 0.25, 1.0, 0.25, 0.25, 0.25, 0.5):
        e, g = model.objective_function(copy.deepcopy(gold['p']), gold['meta']['mb_size'], alpha=alpha)
        gu.assert_close(e, g, {'energy': ref[alpha][0], 'g': ref[alpha][1], 'meta': {}}, 1e-9, 'alpha %g' % alpha)
    assert len(model.sgp_layer._graphs) == 3
