"""Model-level parity on the CPU fiber emulator: the product's host code (geepee_b200 models,
torch fp64 tail) driving the emulated kernels, against golden vectors from the reference."""
import pytest

import emu_util
import golden_util as gu
import model_cases as mc


@pytest.fixture(scope='module', autouse=True)
def emu():
    emu_util.attach()
    yield
    emu_util.detach()


@pytest.mark.parametrize('name', gu.model_cases())
def test_objective_fp64(name):
    mc.check_model(name, 'fp64', 1e-6)


@pytest.mark.parametrize('name', ['aep_sgpr', 'vfe_sgpr', 'aep_sgpr_probit'])
def test_chunked_rows(name):
    mc.check_chunked(name, 'fp64', 1e-6)


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr', 'aep_sgplvm', 'aep_sgpssm_lin', 'vfe_sgpr'])
def test_objective_fp32(name):
    mc.check_model(name, 'fp32', 1e-3)


@pytest.mark.parametrize('name', ['aep_sgpr', 'vfe_sgpr', 'aep_sdgpr', 'aep_sgpr_nonnat', 'aep_sdgprh'])
def test_predict(name):
    mc.check_predict(name, 'fp64', 1e-8)


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr'])
def test_sampling(name):
    mc.check_sampling(name, 1e-5)


@pytest.mark.parametrize('name', ['aep_sgpssm_lin', 'aep_sgpssm_gp'])
def test_ssm_predict(name):
    mc.check_ssm_predict(name, 1e-8)


def test_predict_with_input_grad():
    mc.check_input_grad(1e-8)


def test_layer_interface():
    mc.check_layer_iface(1e-7)


@pytest.mark.parametrize('name', ['vfe_sgpr', 'vfe_sgpr_probit', 'vfe_sgplvm', 'vfe_sgplvm_probit'])
def test_aep_alpha_to_zero_is_vfe(name):
    mc.check_aep_to_vfe_limit(name)


@pytest.mark.parametrize('name', ['aep_sgpr', 'aep_sdgpr', 'aep_sgplvm', 'aep_sgpssm_lin_1d', 'vfe_sgpr',
                                  'vfe_sgplvm', 'aep_sgpr_probit'])
def test_finite_differences(name):
    mc.check_finite_differences(name, per_key=1)     # the fiber emulator is slow; GPU twin checks 3 per key


def test_lik_interface():
    mc.check_lik_iface(1e-9)


def test_gauss_emis_limit():
    mc.check_gauss_emis_limit()


def test_psi_with_zero_variance_is_the_kernel():
    mc.check_psi_zero_variance_is_kernel()
