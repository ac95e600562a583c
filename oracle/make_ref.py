#!/usr/bin/env python
"""oracle/make_ref.py -- TEST INFRASTRUCTURE, not product code.

Builds two things under the git-ignored ``oracle/_ref/``:

1. ``libgeepee_oracle.so`` from ``oracle/psi_oracle.c`` (always; gcc only).
2. ``geepee/`` -- a mechanically patched, Python-3-importable copy of the seven
   reference modules on the hot path, produced from the sources where they lie
   under /root/reference (only when that tree is present, i.e. in the build
   container).  The copy is used HERE to pin ``oracle/geepee_oracle.py`` and to
   generate ``tests/golden/*.npz`` (tests/golden/gen_golden.py); it is never
   committed and nothing on the GPU box needs it.

The patches are non-semantic:
  * py2 ``print x``            -> ``print(x)``
  * ``import cPickle``         -> ``import pickle``
  * ``import __builtin__``     -> ``import builtins as __builtin__``
  * implicit relative imports  -> explicit (``from .config import *`` ...)
  * ``M * (M + 1) / 2``        -> ``//`` (py2 integer division)
  * numpy>=2 changed ``np.linalg.solve(A[d,M,M], b[d,M])`` (b is now a matrix,
    not a stack of vectors): rewritten to ``solve(A, b[..., None])[..., 0]``
  * ``weave`` / ``matplotlib`` are absent: stub modules; ``compute_psi_weave``
    keeps its Python prologue and calls the C restatement of its own inline
    C++ body (oracle/psi_oracle.c) through ctypes instead of weave.inline.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('GEEPEE_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')

MODULES = ['config', 'utils', 'kernels', 'lik_layers', 'base_models',
           'aep_models', 'vfe_models']


def build_c():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, 'libgeepee_oracle.so')
    src = os.path.join(HERE, 'psi_oracle.c')
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        # -O2 like a default weave/distutils build; no -ffast-math (IEEE order)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', so, src, '-lm'])
    return so


PSI_TAIL = '''

# ---- appended by oracle/make_ref.py: weave is not available --------------
def compute_psi_weave(lls2, lsf2, xmean, xvar, z):
    """Same prologue as the original; the inline C++ body is executed from
    oracle/psi_oracle.c (a flat-index restatement of that body) via ctypes."""
    import ctypes
    ls2 = np.exp(lls2)
    sf2 = float(np.exp(lsf2).ravel()[0])
    M = z.shape[0]
    Q = z.shape[1]
    N = xmean.shape[0]
    lsp2xvar = ls2 + 2.0 * xvar
    log_denom_psi2 = np.ascontiguousarray(0.5 * np.log(ls2 / lsp2xvar))
    lspxvar = ls2 + xvar
    log_denom_psi1 = np.ascontiguousarray(0.5 * np.log(ls2 / lspxvar))
    psi2 = np.empty((N, M, M))
    psi1 = np.empty((N, M))
    lib = _oracle_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    ls2c, zc, mc, vc = c(ls2), c(z), c(xmean), c(xvar)
    lib.geepee_oracle_psi(
        ctypes.c_long(N), ctypes.c_long(M), ctypes.c_long(Q), ctypes.c_double(sf2),
        ls2c.ctypes.data_as(dp), zc.ctypes.data_as(dp), mc.ctypes.data_as(dp),
        vc.ctypes.data_as(dp), log_denom_psi1.ctypes.data_as(dp),
        log_denom_psi2.ctypes.data_as(dp), psi1.ctypes.data_as(dp),
        psi2.ctypes.data_as(dp))
    return psi1, psi2


_ORACLE_LIB = None


def _oracle_lib():
    global _ORACLE_LIB
    if _ORACLE_LIB is None:
        import ctypes, os
        here = os.path.dirname(os.path.abspath(__file__))
        _ORACLE_LIB = ctypes.CDLL(os.path.join(here, '..', 'libgeepee_oracle.so'))
        _ORACLE_LIB.geepee_oracle_psi.restype = None
    return _ORACLE_LIB
'''


def patch(name, src):
    # print statements (only the simple one-line forms that occur in these files)
    src = re.sub(r"^(\s*)print (?!\()(.+)$", r"\1print(\2)", src, flags=re.M)
    src = src.replace('import cPickle as pickle', 'import pickle')
    src = src.replace('import __builtin__', 'import builtins as __builtin__')
    for mod in MODULES:
        src = re.sub(r"^(\s*)from %s import" % mod, r"\1from .%s import" % mod,
                     src, flags=re.M)
    src = src.replace('(M + 1) / 2', '(M + 1) // 2')
    src = src.replace('(self.M + 1) / 2', '(self.M + 1) // 2')
    # numpy>=2 batched solve with stacked vectors
    src = src.replace(
        'np.linalg.solve(\n            self.Su, self.mu)',
        'np.linalg.solve(\n            self.Su, self.mu[..., None])[..., 0]')
    src = src.replace(
        'np.linalg.solve(\n            self.Suhat, self.muhat)',
        'np.linalg.solve(\n            self.Suhat, self.muhat[..., None])[..., 0]')
    src = src.replace('VinvY = np.linalg.solve(Vy, Ydiff)',
                      'VinvY = np.linalg.solve(Vy, Ydiff[..., None])[..., 0]')
    if name == 'kernels':
        src += PSI_TAIL
    if name == 'aep_models':
        # SDGPR_H (aep_models.py:1493,1495) names Gauss_Layer / Probit_Layer, which the module never
        # imports (NameError as shipped): add the missing import, nothing else
        src = src.replace('from .base_models import Base_Model\n',
                          'from .base_models import Base_Model\nfrom .lik_layers import Gauss_Layer, Probit_Layer\n', 1)
    return src


def build_py():
    if not os.path.isdir(os.path.join(REF, 'geepee')):
        return None
    pkg = os.path.join(OUT, 'geepee')
    os.makedirs(pkg, exist_ok=True)
    open(os.path.join(pkg, '__init__.py'), 'w').close()
    for mod in MODULES:
        with open(os.path.join(REF, 'geepee', mod + '.py')) as f:
            src = f.read()
        with open(os.path.join(pkg, mod + '.py'), 'w') as f:
            f.write(patch(mod, src))
    # stub modules for the absent imports
    stubs = os.path.join(OUT, 'stubs')
    os.makedirs(os.path.join(stubs, 'matplotlib'), exist_ok=True)
    with open(os.path.join(stubs, 'weave.py'), 'w') as f:
        f.write("class converters:\n    blitz = None\n\n"
                "def inline(*a, **k):\n"
                "    raise RuntimeError('weave is stubbed (oracle/make_ref.py)')\n")
    for m in ['__init__', 'pyplot', 'pylab']:
        open(os.path.join(stubs, 'matplotlib', m + '.py'), 'w').close()
    return pkg


def import_ref():
    """Return (aep_models, vfe_models, lik_layers, kernels, utils) of the patched copy."""
    build_c()
    pkg = build_py()
    if pkg is None:
        raise RuntimeError('reference tree not found at %s' % REF)
    stubs = os.path.join(OUT, 'stubs')
    # stubs go LAST so a real matplotlib, if present, wins
    if stubs not in sys.path:
        sys.path.append(stubs)
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import importlib
    mods = [importlib.import_module('geepee.' + m)
            for m in ['aep_models', 'vfe_models', 'lik_layers', 'kernels', 'utils']]
    return tuple(mods)


if __name__ == '__main__':
    print('built', build_c())
    print('patched copy:', build_py())
