"""predict_f_with_input_grad / predict_y_with_input_grad / backprop_predictive_grads_reg on the B200
(det_fwd + det_dx kernels with the posterior operands) against outputs of the reference itself
(tests/golden/input_grad.npz).  Added after the round's last GPU session: verified on the CPU fiber
emulator only (tests/test_emu_models.py::test_predict_with_input_grad), hence collected last."""
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
