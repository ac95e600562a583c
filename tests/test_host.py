"""Host-side glue (utils): flatten/unflatten, ObjectiveWrapper NaN policy, adam; and the
optimise() loop driving the emulated device path end to end."""
import numpy as np
import pytest

import emu_util
from geepee_b200 import utils


def test_flatten_roundtrip_sorted_keys():
    p = {'zu': np.arange(6.).reshape(3, 2), 'ls': np.array([1., 2.]), 'sn': np.array(0.5), 'sf': np.array([3.])}
    vec, args = utils.flatten_dict(p)
    assert vec.tolist() == [1., 2., 3., 0.5, 0., 1., 2., 3., 4., 5.]      # ls, sf, sn, zu
    q = utils.unflatten_dict(vec, args)
    assert all(np.array_equal(p[k], q[k]) and p[k].shape == q[k].shape for k in p)


def test_objective_wrapper_replaces_nonfinite():
    class Obj(object):
        def objective_function(self, params, idxs, alpha, prop_mode):
            return 1.5, {'a': np.array([1.0, np.nan]), 'b': np.array([np.inf])}
    w = utils.ObjectiveWrapper()
    vec, args = utils.flatten_dict({'a': np.zeros(2), 'b': np.zeros(1)})
    f, g = w(vec, args, Obj(), 3, 0.5, 'MM')
    assert f == 1.5 and g.tolist() == [1.0, 0.0, 0.0]


def test_adam_minimises_quadratic():
    x = utils.adam(lambda x, *a: (float(np.sum(x**2)), 2 * x), np.array([1.0, -2.0]), maxiter=300,
                   step_size=0.05, args=(), disp=False)
    assert np.all(np.abs(x) < 0.05)


def test_optimise_reduces_energy_on_emulator():
    emu_util.attach()
    try:
        from geepee_b200 import aep_models as aep
        rng = np.random.RandomState(0)
        x = rng.rand(40, 1)
        y = np.sin(6 * x) + 0.1 * rng.standard_normal((40, 1))
        np.random.seed(0)
        model = aep.SGPR(x, y, 6)
        p0 = model.init_hypers(y)
        e0, _ = model.objective_function(p0, 40, alpha=0.5)
        np.random.seed(0)
        model.optimise(method='L-BFGS-B', alpha=0.5, maxiter=15, disp=False)
        e1, _ = model.objective_function(model.get_hypers(), 40, alpha=0.5)
        assert e1 < e0
        mf, vf = model.predict_f(x[:5])
        assert mf.shape == (5, 1) and np.all(vf > 0)
    finally:
        emu_util.detach()


def test_nvtx_annotate_is_identity_when_off_and_wraps_when_on(monkeypatch):
    """geepee_b200/nvtx.py: no wrapper on the default path; with GPB_NVTX=1 the phase runs between a push and a pop
    (also when it raises)."""
    import torch
    from geepee_b200 import nvtx

    def f(a, b=1):
        return a + b

    monkeypatch.setattr(nvtx, 'ENABLED', False)
    assert nvtx.annotate('x')(f) is f
    calls = []
    monkeypatch.setattr(nvtx, 'ENABLED', True)
    monkeypatch.setattr(torch.cuda.nvtx, 'range_push', lambda s: calls.append(('push', s)))
    monkeypatch.setattr(torch.cuda.nvtx, 'range_pop', lambda: calls.append(('pop',)))
    g = nvtx.annotate('phase')(f)
    assert g(2, b=3) == 5 and calls == [('push', 'geepee/phase'), ('pop',)]

    def boom():
        raise ValueError('x')
    with pytest.raises(ValueError):
        nvtx.annotate('boom')(boom)()
    assert calls[-1] == ('pop',) and len(calls) == 4
