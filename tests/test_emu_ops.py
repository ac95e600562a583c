"""Kernel-level parity on the CPU fiber emulator (tests/emu): the SAME kernel and launch
source as the CUDA library, compiled for the host, checked against the oracle's numpy
formulas.  The `-m gpu` twin of this file is tests/test_gpu_ops.py."""
import pytest

import emu_util
import ops_cases as oc


@pytest.fixture(scope='module', autouse=True)
def emu():
    emu_util.attach()
    oc.DEV = 'cpu'
    yield
    emu_util.detach()


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-12), ('fp32', 2e-4)])
@pytest.mark.parametrize('n,M,D,Do', oc.DET_SHAPES)
def test_det_layer(n, M, D, Do, prec, tol):
    oc.check_det_layer(n, M, D, Do, prec, tol)


@pytest.mark.parametrize('prec,tol', [('fp64', 1e-12), ('fp32', 3e-4)])
@pytest.mark.parametrize('n,M,Q,Do', oc.MM_SHAPES)
def test_mm_layer(n, M, Q, Do, prec, tol):
    oc.check_mm_layer(n, M, Q, Do, prec, tol)


def test_kmat_psi_lik():
    oc.check_kmat_psi_lik()


def test_torch_custom_ops():
    oc.check_torch_custom_ops()


@pytest.mark.parametrize('n,Do,Q', oc.EMIS_SHAPES)
def test_gauss_emis(n, Do, Q):
    oc.check_gauss_emis(n, Do, Q)


def test_probit_lik():
    oc.check_probit_lik()


@pytest.mark.parametrize('M,batch', [(5, 1), (32, 2), (50, 3), (77, 1)])
def test_spd_inverse(M, batch):
    oc.check_spd_inverse(M, batch)
