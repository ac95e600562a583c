"""Optimiser glue kept on the host, with the reference's semantics (geepee/utils.py):
ObjectiveWrapper (37-53), flatten_dict / unflatten_dict (68-90), adam (93-117),
PCA_reduce (21-34).  These are the callers of the hot path, not part of it."""
import numpy as np


def profile(func):
    """geepee/utils.py:6-10: pass-through when no line profiler is installed."""
    return func


def PCA_reduce(X, Q):
    """geepee/utils.py:21-34 (kept bug-for-bug: eigh's outputs are named the other way round)."""
    assert Q <= X.shape[1], 'Cannot have more latent dimensions than observed'
    evecs, evals = np.linalg.eigh(np.cov(X.T))
    i = np.argsort(evecs)[::-1]
    W = evals[:, i]
    W = W[:, :Q]
    return (X - X.mean(0)).dot(W)


def flatten_dict(params):
    """geepee/utils.py:68-82: concatenate values in sorted-key order."""
    keys = list(params.keys())
    shapes = {}
    sizes = []
    chunks = []
    for key in sorted(keys):
        val = np.asarray(params[key])
        shapes[key] = val.shape
        chunks.append(val.ravel())
        sizes.append(val.size)
    vec = np.concatenate(chunks) if chunks else np.array([])
    indices = np.cumsum(np.array(sizes, dtype=int))[:-1]
    return vec, (keys, indices, shapes)


def unflatten_dict(params, params_args):
    """geepee/utils.py:85-90."""
    keys, indices, shapes = params_args[0], params_args[1], params_args[2]
    vals = np.split(params, indices)
    return {key: np.reshape(vals[i], shapes[key]) for i, key in enumerate(sorted(keys))}


class ObjectiveWrapper(object):
    """geepee/utils.py:37-53: vector <-> dict adapter; non-finite gradient entries are
    replaced by zeros with a warning."""

    def __init__(self):
        self.previous_x = None

    def __call__(self, params, params_args, obj, idxs, alpha, prop_mode):
        params_dict = unflatten_dict(params, params_args)
        f, grad_dict = obj.objective_function(params_dict, idxs, alpha=alpha, prop_mode=prop_mode)
        g, _ = flatten_dict(grad_dict)
        f = float(np.ravel(f)[0])
        fin = np.isfinite(g)
        if np.all(fin):
            self.previous_x = params
            return f, g
        print("Warning: inf or nan in gradient: replacing with zeros")
        return f, np.where(fin, g, 0.)


def adam(func, init_params, callback=None, maxiter=1000, step_size=0.001, b1=0.9, b2=0.999,
         eps=1e-8, args=None, disp=True, return_cost=False):
    """geepee/utils.py:93-117 (Adam, arXiv:1412.6980)."""
    x = init_params
    m = np.zeros_like(x)
    v = np.zeros_like(x)
    fs = []
    for i in range(maxiter):
        f, g = func(x, *args)
        if disp and i % 10 == 0:
            print('iter %d \t obj %.3f' % (i, f))
        if callback:
            callback(x, i, args)
        m = (1 - b1) * g + b1 * m
        v = (1 - b2) * (g**2) + b2 * v
        mhat = m / (1 - b1**(i + 1))
        vhat = v / (1 - b2**(i + 1))
        x = x - step_size * mhat / (np.sqrt(vhat) + eps)
        fs.append(f)
    if return_cost:
        return x, np.array(fs)
    return x
