"""Kernel-level parity on the B200: every C-ABI op of libgeepee_b200.so against the oracle's
numpy formulas (same bodies as tests/test_emu_ops.py), at small shapes plus shapes that span
several tiles / blocks / row splits."""
import pytest
import torch

import ops_cases as oc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()          # raises if libgeepee_b200.so is missing: no fallback
    oc.DEV = 'cuda'
    yield
    oc.DEV = 'cpu'


GPU_DET = oc.DET_SHAPES + [(5000, 200, 10, 2), (3000, 512, 16, 1), (20000, 50, 1, 1)]
GPU_MM = oc.MM_SHAPES + [(700, 64, 5, 3), (2000, 40, 2, 2), (300, 128, 4, 5)]


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-11), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,D,Do', GPU_DET)
def test_det_layer(n, M, D, Do, prec, tol):
    oc.check_det_layer(n, M, D, Do, prec, tol)


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-11), ('fp32', 5e-4)])
@pytest.mark.parametrize('n,M,Q,Do', GPU_MM)
def test_mm_layer(n, M, Q, Do, prec, tol):
    oc.check_mm_layer(n, M, Q, Do, prec, tol)


def test_kmat_psi_lik():
    oc.check_kmat_psi_lik()


@pytest.mark.parametrize('n,Do,Q', oc.EMIS_SHAPES)
def test_gauss_emis(n, Do, Q):
    oc.check_gauss_emis(n, Do, Q)


def test_probit_lik():
    oc.check_probit_lik()


@pytest.mark.parametrize('M,batch', [(5, 1), (32, 2), (50, 3), (77, 1), (128, 4), (200, 2), (256, 5), (512, 2)])
def test_spd_inverse(M, batch):
    oc.check_spd_inverse(M, batch)
