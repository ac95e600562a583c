"""bench.py's end-to-end input path (InputPipeline: double-buffered upload of the step's rows on a copy stream, the
parameter vector read from mapped pinned memory by a copy kernel) must be invisible: every step sees exactly the rows
that were in the pinned host buffers when its upload was issued, and the energy / gradients equal those of the
resident-data call on the same rows."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()
    yield


def _model(N, D, M, seed):
    from geepee_b200 import aep_models as aep
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((N, D))
    Y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, 1))
    np.random.seed(seed)
    model = aep.SGPR(X, Y, M, device=torch.device('cuda'))
    params = model.init_hypers(Y)
    return model, params, X, Y


def test_pipelined_steps_match_resident_steps():
    import bench
    N, D, M = 60000, 10, 32                     # 4.8 MB of x per step: above the pipeline's 4 MB threshold
    model, params, X, Y = _model(N, D, M, 0)
    dev = torch.device('cuda')
    # reference answers with resident data, for the original rows and for a second data set
    e_a, g_a = model.objective_function(params, N, alpha=0.5)
    rng = np.random.RandomState(1)
    Xb = X + 0.3 * rng.standard_normal(X.shape)
    Yb = Y + 0.3 * rng.standard_normal(Y.shape)
    ref = _model(N, D, M, 0)[0]
    ref._x.copy_(torch.from_numpy(Xb))
    ref._y.copy_(torch.from_numpy(Yb))
    e_b, g_b = ref.objective_function(params, N, alpha=0.5)
    assert abs(e_a - e_b) > 1e-6 * abs(e_a)     # the two data sets are distinguishable

    xh, yh = torch.from_numpy(X.copy()).pin_memory(), torch.from_numpy(Y.copy()).pin_memory()
    pipe = bench.InputPipeline(model, xh, yh, 0, N, dev)
    assert pipe.pipelined

    def step():
        return pipe.step(lambda: model.objective_function(params, N, alpha=0.5))

    def same(e, g, e0, g0):
        assert abs(e - e0) <= 1e-9 * abs(e0)
        for k in g0:
            assert np.max(np.abs(g[k] - g0[k])) <= 1e-8 * max(1e-30, np.max(np.abs(g0[k]))), k

    # step 0 uploads A for itself and prefetches A for step 1
    same(*step(), e_a, g_a)
    torch.cuda.synchronize()                    # step 1's prefetch (data set A) has landed
    xh.copy_(torch.from_numpy(Xb))              # host buffers now hold data set B
    yh.copy_(torch.from_numpy(Yb))
    same(*step(), e_a, g_a)                     # step 1 still computes on A (uploaded before the change) ...
    same(*step(), e_b, g_b)                     # ... step 2 on B (its upload was issued during step 1)
    same(*step(), e_b, g_b)
    pipe.finish()
    torch.cuda.synchronize()
    # the two device buffers alternate and neither is the other's alias
    assert pipe.bufs[0][0].data_ptr() != pipe.bufs[1][0].data_ptr()


def test_small_inputs_use_the_in_stream_copy():
    import bench
    N, D, M = 200, 1, 10
    model, params, X, Y = _model(N, D, M, 3)
    e0, g0 = model.objective_function(params, N, alpha=0.5)
    xh, yh = torch.from_numpy(X.copy()).pin_memory(), torch.from_numpy(Y.copy()).pin_memory()
    pipe = bench.InputPipeline(model, xh, yh, 0, N, torch.device('cuda'))
    assert not pipe.pipelined
    e, g = pipe.step(lambda: model.objective_function(params, N, alpha=0.5))
    pipe.finish()
    assert abs(e - e0) <= 1e-9 * abs(e0)


def test_zero_copy_parameter_upload_matches_dma(monkeypatch):
    """layers._zero_copy_upload (gpb_host_device_ptr + gpb_tail_copy reading mapped pinned memory) against the
    copy-engine upload and the host values."""
    from geepee_b200 import layers
    model, params, X, Y = _model(500, 3, 16, 5)
    dev = torch.device('cuda')
    monkeypatch.delenv('GPB_UPLOAD_DMA', raising=False)
    a = layers.pack_to_device(params, dev)
    torch.cuda.synchronize()
    assert any(v != 0 for v in layers._ZC.values()), 'the zero-copy path was not taken'
    monkeypatch.setenv('GPB_UPLOAD_DMA', '1')
    b = layers.pack_to_device(params, dev)
    torch.cuda.synchronize()
    assert sorted(a) == sorted(b) == sorted(params)
    for k in params:
        assert torch.equal(a[k], b[k]), k
        assert np.array_equal(a[k].cpu().numpy().reshape(-1), np.asarray(params[k], dtype=np.float64).reshape(-1)), k
