"""GPU twins of tests added after the round's last GPU session (GPU budget spent): input gradients
of the prediction (det_fwd + det_dx kernels with the posterior operands), the layer-level interface
of SURVEY 8b, the AEP(alpha -> 0) = VFE identity and the reference's finite-difference harness, all
through the product.  Each has been verified on the CPU fiber emulator (tests/test_emu_models.py);
the file name makes pytest collect it last, after the GPU-verified suites."""
import pytest
import torch

import model_cases as mc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_lib():
    from geepee_b200 import _lib
    _lib._testing_detach()
    assert torch.cuda.is_available()
    _lib.get()
    yield


def test_predict_with_input_grad():
    mc.check_input_grad(1e-7)


def test_layer_interface():
    """SURVEY 8b layer-level interface (tests/golden/layer_iface.npz); emulator twin:
    tests/test_emu_models.py::test_layer_interface."""
    mc.check_layer_iface(1e-6)


@pytest.mark.parametrize('name', ['vfe_sgpr', 'vfe_sgpr_probit', 'vfe_sgplvm', 'vfe_sgplvm_probit'])
def test_aep_alpha_to_zero_is_vfe(name):
    """tests/test_aep_vfe_limits.py:17-127 on the B200; emulator twin in tests/test_emu_models.py."""
    mc.check_aep_to_vfe_limit(name)


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr', 'aep_sgplvm', 'aep_sgpssm_lin_1d', 'vfe_sgpr',
                                  'vfe_sgplvm', 'aep_sgpr_probit'])
def test_finite_differences(name):
    """The reference's tests/test_grads_* harness (tests/test_utils.py:61-138) on the B200."""
    mc.check_finite_differences(name, per_key=3)


def test_lik_interface():
    """Public Gauss_Layer / Probit_Layer interface incl. the Monte-Carlo 3-D branches."""
    mc.check_lik_iface(1e-8)


def test_gauss_emis_limit():
    """tests/test_grads_emis.py:212-240 on the B200."""
    mc.check_gauss_emis_limit()


def test_psi_with_zero_variance_is_the_kernel():
    """TODO.txt:65-66 of the reference (SURVEY 8c pin iv) on the B200."""
    mc.check_psi_zero_variance_is_kernel()
