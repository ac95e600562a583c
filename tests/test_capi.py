"""The C-ABI library: every symbol declared in include/geepee_b200.h is exported by the CUDA
build (dlopen + dlsym only -- no compute call, this runs without a GPU) and by the emulator
build, and the product refuses to run without CUDA."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'geepee_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gpb_[a-z0-9_]+)\s*\(', src)))


def test_header_and_binding_table_agree():
    from geepee_b200 import _lib
    assert declared_symbols() == _lib.EXPORTS


def test_cuda_library_exports_every_symbol():
    from geepee_b200 import _lib, build
    build.build()                      # no-op when libgeepee_b200.so is up to date
    lib = _lib.load_library()          # binds every symbol; AttributeError if one is missing
    for name in declared_symbols():
        assert hasattr(lib, name)
    assert lib.gpb_version() >= 100
    assert lib.gpb_det_pad_m(200) == 256 and lib.gpb_det_pad_m(513) == -1


def test_emulator_library_exports_every_symbol():
    import emu_util
    lib = emu_util.attach()
    try:
        for name in declared_symbols():
            assert hasattr(lib.get(), name)
    finally:
        emu_util.detach()


def test_no_cpu_fallback():
    import torch
    from geepee_b200 import _lib
    _lib._testing_detach()
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.get()
    import numpy as np
    from geepee_b200 import aep_models
    with pytest.raises(RuntimeError):
        m = aep_models.SGPR(np.zeros((4, 1)), np.zeros((4, 1)), 2)
        m.objective_function(m.init_hypers(np.zeros((4, 1))), 4)
