"""Loader of the C-ABI shared library (include/geepee_b200.h).

The product path is CUDA only: if ``geepee_b200/csrc/libgeepee_b200.so`` is missing,
or no CUDA device is visible, every op raises -- there is no CPU fallback.
``_testing_attach`` exists for the repository's own CPU test-suite, which loads the
fiber-emulated twin of the library (tests/emu) to exercise kernel and host logic in the
GPU-less build container; nothing in the package ever calls it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPB_LIB_PATH: another build of the same CUDA library (development A/B variants, build.py)
LIB_PATH = os.environ.get('GPB_LIB_PATH') or os.path.join(_HERE, 'csrc', 'libgeepee_b200.so')

_lib = None
_device_type = 'cuda'

c_dp = ctypes.c_void_p

_SIGS = {
    'gpb_version': (ctypes.c_int, []),
    'gpb_last_error': (ctypes.c_char_p, []),
    'gpb_sm_count': (ctypes.c_int, []),
    'gpb_launch_count': (ctypes.c_long, []),
    'gpb_prec_bytes': (ctypes.c_int, [ctypes.c_int]),
    'gpb_kmat': (ctypes.c_int, [c_dp] * 4 + [ctypes.c_int] * 3 + [ctypes.c_double, c_dp, c_dp]),
    'gpb_psi_stats': (ctypes.c_int, [c_dp] * 5 + [ctypes.c_int] * 3 + [c_dp, c_dp, c_dp]),
    'gpb_gauss_lik_ws_bytes': (ctypes.c_size_t, [ctypes.c_long]),
    'gpb_gauss_lik': (ctypes.c_int, [c_dp] * 4 + [ctypes.c_double, ctypes.c_double, ctypes.c_long,
                                                  ctypes.c_int, c_dp, c_dp, c_dp, c_dp,
                                                  ctypes.c_size_t, c_dp]),
    'gpb_det_pad_m': (ctypes.c_int, [ctypes.c_int]),
    'gpb_det_pad_operands': (ctypes.c_int, [ctypes.c_int, c_dp, c_dp, ctypes.c_int, ctypes.c_int,
                                            c_dp, c_dp, c_dp]),
    'gpb_det_fwd': (ctypes.c_int, [ctypes.c_int] + [c_dp] * 6 + [ctypes.c_int] * 4 + [c_dp] * 5),
    'gpb_det_tc_available': (ctypes.c_int, []),
    'gpb_det_tc_bu_bytes': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'gpb_det_tc_zs_bytes': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'gpb_det_tc_prep': (ctypes.c_int, [c_dp, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp]),
    'gpb_det_fwd_tc': (ctypes.c_int, [c_dp] * 6 + [ctypes.c_int] * 4 + [c_dp] * 5),
    'gpb_spd_inverse': (ctypes.c_int, [c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp]),
    'gpb_probit_lik': (ctypes.c_int, [c_dp] * 5 + [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_long,
                                               ctypes.c_int] + [c_dp, c_dp, c_dp, c_dp, ctypes.c_size_t, c_dp]),
    'gpb_gauss_emis_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 3),
    'gpb_gauss_emis': (ctypes.c_int, [c_dp] * 5 + [ctypes.c_double, ctypes.c_double] + [ctypes.c_int] * 3 +
                       [c_dp, c_dp, c_dp, c_dp, ctypes.c_size_t, c_dp]),
    'gpb_gauss_emis_finish': (ctypes.c_int, [c_dp, c_dp, ctypes.c_double, ctypes.c_double, ctypes.c_long, ctypes.c_int,
                                             ctypes.c_int, c_dp, c_dp]),
    'gpb_det_bwd_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 4),
    'gpb_det_bwd': (ctypes.c_int, [ctypes.c_int] + [c_dp] * 9 + [ctypes.c_int] * 4 + [c_dp] * 5 +
                    [ctypes.c_size_t, c_dp]),
    'gpb_det_dx': (ctypes.c_int, [ctypes.c_int] + [c_dp] * 8 + [ctypes.c_int] * 4 + [c_dp, c_dp]),
    'gpb_det_syrk_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 3),
    'gpb_det_syrk': (ctypes.c_int, [ctypes.c_int, c_dp, c_dp] + [ctypes.c_int] * 3 +
                     [c_dp, c_dp, ctypes.c_size_t, c_dp]),
    'gpb_mm_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 5),
    'gpb_mm_fwd': (ctypes.c_int, [ctypes.c_int] + [c_dp] * 7 + [ctypes.c_int] * 4 +
                   [c_dp, c_dp, c_dp, c_dp, c_dp, ctypes.c_size_t, c_dp]),
    'gpb_mm_bwd': (ctypes.c_int, [ctypes.c_int] + [c_dp] * 12 + [ctypes.c_int] * 4 + [c_dp] * 9 +
                   [ctypes.c_size_t, c_dp]),
    'gpb_tail_exec': (ctypes.c_int, [c_dp, ctypes.c_int, c_dp]),
    'gpb_tail_khyper_out_len': (ctypes.c_long, [ctypes.c_int, ctypes.c_int]),
    'gpb_tail_gather': (ctypes.c_int, [ctypes.c_int, c_dp, c_dp, ctypes.c_double, c_dp, c_dp]),
    'gpb_tail_copy': (ctypes.c_int, [ctypes.c_int, c_dp, c_dp, c_dp, c_dp]),
    'gpb_host_device_ptr': (ctypes.c_int, [c_dp, c_dp]),
    'gpb_latent_ws_bytes': (ctypes.c_size_t, [ctypes.c_long]),
    'gpb_lvm_x_fwd': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp, ctypes.c_long, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_dp, c_dp, c_dp]),
    'gpb_lvm_x_bwd': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_dp, c_dp, c_dp, ctypes.c_long, ctypes.c_int,
                                     ctypes.c_long, ctypes.c_int] + [ctypes.c_double] * 5 + [c_dp] * 6 +
                      [ctypes.c_size_t, c_dp]),
    'gpb_ssm_cavity': (ctypes.c_int, [c_dp, c_dp, ctypes.c_long, ctypes.c_int] + [ctypes.c_double] * 3 + [c_dp] * 3),
    'gpb_ssm_transition': (ctypes.c_int, [c_dp] * 5 + [ctypes.c_long, ctypes.c_double, ctypes.c_double] + [c_dp] * 4 +
                           [ctypes.c_size_t, c_dp]),
    'gpb_ssm_sources': (ctypes.c_int, [c_dp, c_dp, ctypes.c_long, ctypes.c_int] + [ctypes.c_double] * 3 +
                        [c_dp, c_dp, ctypes.c_long, ctypes.c_long, ctypes.c_int] * 3 + [c_dp] * 3),
    'gpb_ssm_xfinal': (ctypes.c_int, [c_dp, c_dp, ctypes.c_long, ctypes.c_int] + [ctypes.c_double] * 3 + [c_dp] * 6 +
                       [ctypes.c_size_t, c_dp]),
    'gpb_profile_enable': (ctypes.c_int, [ctypes.c_int]),
    'gpb_profile_collect': (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_long)]),
    'gpb_fma_peak': (ctypes.c_int, [ctypes.c_int, ctypes.c_long, c_dp, ctypes.POINTER(ctypes.c_double), c_dp]),
}

EXPORTS = sorted(_SIGS)


def _bind(lib):
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


def load_library(path=LIB_PATH):
    """dlopen + bind every symbol of include/geepee_b200.h (no CUDA call is made)."""
    if not os.path.exists(path):
        raise RuntimeError(
            'geepee_b200: %s is missing -- build it with `python __graft_entry__.py build` '
            '(nvcc, sm_100a).  There is no CPU fallback.' % path)
    return _bind(ctypes.CDLL(path))


def get():
    global _lib
    if _lib is None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('geepee_b200: no CUDA device visible; the hot path is CUDA only '
                               '(no CPU fallback)')
        _lib = load_library()
    return _lib


def device_type():
    return _device_type


def _testing_attach(path, device_type='cpu'):
    """TEST HOOK (tests/ only): use another build of the same C ABI, e.g. the CPU
    emulation twin from tests/emu.  Never called by the package itself."""
    global _lib, _device_type
    _lib = _bind(ctypes.CDLL(path))
    _device_type = device_type
    return _lib


def _testing_detach():
    global _lib, _device_type
    _lib = None
    _device_type = 'cuda'
