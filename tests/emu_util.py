"""Attach the CPU-emulated twin of libgeepee_b200.so for the GPU-less test-suite."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'emu'))


def attach():
    import build as emu_build
    from geepee_b200 import _lib
    so = emu_build.build()
    _lib._testing_attach(so, 'cpu')
    return _lib


def detach():
    from geepee_b200 import _lib
    _lib._testing_detach()
