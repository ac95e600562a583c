"""Thin torch-tensor wrappers over the C ABI (include/geepee_b200.h).

torch is used for device memory and the current CUDA stream only; every op below is one
call into libgeepee_b200.so with raw device pointers.  Inputs are fp64, C-contiguous
tensors that already live on the device; nothing here copies to the host or synchronises.
"""
import ctypes

import torch

from . import _lib, config

F64 = 0
F32 = 1

PREC = {'fp64': F64, 'f64': F64, 'fp32': F32, 'f32': F32, 'fp32_psi': F32, F64: F64, F32: F32}


def prec_dtype(prec):
    return torch.float64 if prec == F64 else torch.float32


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream(t):
    """torch's current stream on the tensor's device as a raw handle (every launch goes there).  The raw getter
    costs a fraction of a microsecond; torch.cuda.current_stream() builds a Stream object (7-8 us, 15 times per step
    of the small configs)."""
    if t.is_cuda:
        if _raw_stream is not None:
            return ctypes.c_void_p(_raw_stream(t.device.index if t.device.index is not None
                                               else torch.cuda.current_device()))
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return None


def _chk(rc, what):
    if rc != 0:
        msg = _lib.get().gpb_last_error()
        raise RuntimeError('geepee_b200.%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))


def _c(t, dtype=torch.float64):
    """Contract check: right device type, dtype, contiguous."""
    if t.device.type != _lib.device_type():
        raise RuntimeError('geepee_b200: tensor on %s, library expects %s (no CPU fallback)'
                           % (t.device.type, _lib.device_type()))
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError('geepee_b200: expected contiguous %s tensor, got %s%s'
                           % (dtype, t.dtype, '' if t.is_contiguous() else ' (strided)'))
    return t


def _ws(nbytes, like):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=like.device)


_launch_adjust = 0


def launch_count():
    """Kernels of libgeepee_b200.so launched so far: the library's own counter plus the library
    kernels executed through CUDA-graph replays (tailgraph.py; a replay does not pass through
    the C ABI, so it is accounted here with the count seen at capture time)."""
    return int(_lib.get().gpb_launch_count()) + _launch_adjust


def adjust_launch_count(n):
    global _launch_adjust
    _launch_adjust += int(n)


def kmat(x, z, ls, sf, jitter=0.0):
    """kernels.py:10-22 compute_kernel(2*ls, 2*sf, x, z) (+ jitter*I when x is z)."""
    lib = _lib.get()
    n, D = x.shape
    M = z.shape[0]
    out = torch.empty((n, M), dtype=torch.float64, device=x.device)
    _chk(lib.gpb_kmat(_p(_c(x)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), n, M, D, float(jitter),
                      _p(out), _stream(x)), 'kmat')
    return out


def psi_stats(mx, vx, z, ls, sf):
    """kernels.py:181-240 compute_psi_weave(2*ls, 2*sf, mx, vx, z), materialised."""
    lib = _lib.get()
    n, Q = mx.shape
    M = z.shape[0]
    psi1 = torch.empty((n, M), dtype=torch.float64, device=mx.device)
    psi2 = torch.empty((n, M, M), dtype=torch.float64, device=mx.device)
    _chk(lib.gpb_psi_stats(_p(_c(mx)), _p(_c(vx)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), n, M, Q,
                           _p(psi1), _p(psi2), _stream(mx)), 'psi_stats')
    return psi1, psi2


def gauss_lik(m, v, y, sn, alpha, scale, mode):
    """lik_layers.py:104-133 (mode 0) / 183-199 (mode 1).  Returns scaled dm, dv and a
    device tensor [sum of log terms, dsn-sum]."""
    lib = _lib.get()
    total = m.numel()
    dm = torch.empty_like(m)
    dv = torch.empty_like(m)
    out2 = torch.empty(2, dtype=torch.float64, device=m.device)
    nb = lib.gpb_gauss_lik_ws_bytes(total)
    ws = _ws(nb, m)
    _chk(lib.gpb_gauss_lik(_p(_c(m)), _p(_c(v)), _p(_c(y)), _p(_c(sn)), float(alpha), float(scale),
                           total, int(mode), _p(dm), _p(dv), _p(out2), _p(ws), ws.numel(),
                           _stream(m)), 'gauss_lik')
    return dm, dv, out2


def spd_inverse(A):
    """Inverse and log-determinant of a batch of SPD matrices [b,M,M] (or one [M,M]), fp64, with the
    library's cluster Gauss-Jordan kernel (base_models.py:464,471,476; aep_models.py:68,78,91,525,533)."""
    lib = _lib.get()
    single = A.dim() == 2
    Ab = _c(A.reshape(-1, A.shape[-1], A.shape[-1]))
    b, M = Ab.shape[0], Ab.shape[-1]
    inv = torch.empty_like(Ab)
    ld = torch.empty(b, dtype=torch.float64, device=A.device)
    _chk(lib.gpb_spd_inverse(_p(Ab), b, M, _p(inv), _p(ld), _stream(A)), 'spd_inverse')
    return (inv[0], ld[0]) if single else (inv, ld)


def probit_lik(m, v, y, gh_x, gh_w, alpha, scale, mode):
    """lik_layers.py:303-362 (mode 0) / 418-436 (mode 1).  Returns scaled dm, dv and a device
    tensor [sum of log terms, 0]."""
    lib = _lib.get()
    total = m.numel()
    dm = torch.empty_like(m)
    dv = torch.empty_like(m)
    out2 = torch.empty(2, dtype=torch.float64, device=m.device)
    ws = _ws(lib.gpb_gauss_lik_ws_bytes(total), m)
    _chk(lib.gpb_probit_lik(_p(_c(m)), _p(_c(v)), _p(_c(y)), _p(_c(gh_x)), _p(_c(gh_w)), int(gh_x.numel()),
                            float(alpha), float(scale), total, int(mode), _p(dm), _p(dv), _p(out2), _p(ws),
                            ws.numel(), _stream(m)), 'probit_lik')
    return dm, dv, out2


def gauss_emis_supported(Do, Q):
    return Do <= 8 and Q <= 8


def gauss_emis(mx, vx, y, C, R, alpha, scale):
    """lik_layers.py:573-627 (tilted linear-Gaussian emission), per-row part.  R: variances.
    -> scale*dmx, scale*dvx, and the unscaled sums [quad | logdet | dRacc[Do] | dC[Do*Q]]."""
    lib = _lib.get()
    n, Q = mx.shape
    Do = y.shape[1]
    dmx = torch.empty_like(mx)
    dvx = torch.empty_like(mx)
    out = torch.empty(2 + Do + Do * Q, dtype=torch.float64, device=mx.device)
    ws = _ws(lib.gpb_gauss_emis_ws_bytes(n, Do, Q), mx)
    _chk(lib.gpb_gauss_emis(_p(_c(mx)), _p(_c(vx)), _p(_c(y)), _p(_c(C)), _p(_c(R)), float(alpha), float(scale),
                            n, Q, Do, _p(dmx), _p(dvx), _p(out), _p(ws), ws.numel(), _stream(mx)), 'gauss_emis')
    return dmx, dvx, out


def gauss_emis_finish(raw, R, alpha, scale, Nb, Do, Q):
    """lik_layers.py:600-627 from the sums of gauss_emis -> [scale*logZ | 0 | scale*dR[Do] | scale*dC[Do*Q]]."""
    fin = torch.empty_like(raw)
    _chk(_lib.get().gpb_gauss_emis_finish(_p(_c(raw)), _p(_c(R)), float(alpha), float(scale), int(Nb), int(Do), int(Q),
                                          _p(fin), _stream(raw)), 'gauss_emis_finish')
    return fin


class DetOperands(object):
    """Zero-padded, precision-typed copies of (A, B_det) for the deterministic layer."""

    def __init__(self, prec, A, B):
        lib = _lib.get()
        self.prec = prec
        self.Do, self.M = A.shape
        self.MP = lib.gpb_det_pad_m(self.M)
        if self.MP < 0:
            raise RuntimeError('geepee_b200: M=%d unsupported by the deterministic layer kernels (max 512)' % self.M)
        dt = prec_dtype(prec)
        self.Ap = torch.empty((self.Do, self.MP), dtype=dt, device=A.device)
        self.Bp = torch.empty((self.Do, self.MP, self.MP), dtype=dt, device=A.device)
        _chk(lib.gpb_det_pad_operands(prec, _p(_c(A)), _p(_c(B)), self.M, self.Do, _p(self.Ap),
                                      _p(self.Bp), _stream(A)), 'det_pad_operands')


def det_fwd(prec, x, z, ls, sf, opnd, save=True):
    """aep_models.py:142-158 / base_models.py:265-284.  Returns mout, vout and (if save)
    the on-device Kfu[n,MP] and T[n,Do,MP] buffers the backward kernels stream."""
    lib = _lib.get()
    n, D = x.shape
    Do, M, MP = opnd.Do, opnd.M, opnd.MP
    mout = torch.empty((n, Do), dtype=torch.float64, device=x.device)
    vout = torch.empty((n, Do), dtype=torch.float64, device=x.device)
    Ks = Ts = None
    if save:
        dt = prec_dtype(prec)
        Ks = torch.empty((n, MP), dtype=dt, device=x.device)
        Ts = torch.empty((n, Do, MP), dtype=dt, device=x.device)
    if prec == F32 and save and config.DET_FP32_TENSOR_CORES and lib.gpb_det_tc_available():
        # fp32-psi mode on tcgen05: 3xTF32 Kfu . B_d, accumulators in TMEM (csrc/gpb_umma.cuh)
        Bu = torch.empty(lib.gpb_det_tc_bu_bytes(M, Do) // 4, dtype=torch.float32, device=x.device)
        Zs = torch.empty(lib.gpb_det_tc_zs_bytes(M, D) // 4, dtype=torch.float32, device=x.device)
        _chk(lib.gpb_det_tc_prep(_p(opnd.Bp), _p(_c(z)), _p(_c(ls)), M, D, Do, _p(Bu), _p(Zs), _stream(x)), 'det_tc_prep')
        _chk(lib.gpb_det_fwd_tc(_p(_c(x)), _p(_c(ls)), _p(_c(sf)), _p(Zs), _p(opnd.Ap), _p(Bu), n, M, D, Do,
                                _p(mout), _p(vout), _p(Ks), _p(Ts), _stream(x)), 'det_fwd_tc')
        return mout, vout, Ks, Ts
    _chk(lib.gpb_det_fwd(prec, _p(_c(x)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), _p(opnd.Ap), _p(opnd.Bp),
                         n, M, D, Do, _p(mout), _p(vout), _p(Ks), _p(Ts), _stream(x)), 'det_fwd')
    return mout, vout, Ks, Ts


def det_bwd(prec, x, z, ls, sf, opnd, dm, dv, Ks, Ts):
    """aep_models.py:452-460,490 + kernels.py:381-399.  -> dA, dzu, dl, dsf2 (device)."""
    lib = _lib.get()
    n, D = x.shape
    Do, M = opnd.Do, opnd.M
    dev = x.device
    dA = torch.empty((Do, M), dtype=torch.float64, device=dev)
    dzu = torch.empty((M, D), dtype=torch.float64, device=dev)
    dl = torch.empty((D,), dtype=torch.float64, device=dev)
    dsf2 = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = _ws(lib.gpb_det_bwd_ws_bytes(n, M, D, Do), x)
    _chk(lib.gpb_det_bwd(prec, _p(_c(x)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), _p(opnd.Ap), _p(_c(dm)),
                         _p(_c(dv)), _p(Ks), _p(Ts), n, M, D, Do, _p(dA), _p(dzu), _p(dl), _p(dsf2),
                         _p(ws), ws.numel(), _stream(x)), 'det_bwd')
    return dA, dzu, dl, dsf2


def det_dx(prec, x, z, ls, opnd, dm, dv, Ks, Ts):
    """aep_models.py:346-350 + kernels.py:393-395: gradient wrt the inputs of the deterministic
    layer (Monte-Carlo propagation feeds it samples of an uncertain input).  -> dx[n,D]."""
    lib = _lib.get()
    n, D = x.shape
    dx = torch.empty((n, D), dtype=torch.float64, device=x.device)
    _chk(lib.gpb_det_dx(prec, _p(_c(x)), _p(_c(z)), _p(_c(ls)), _p(opnd.Ap), _p(_c(dm)), _p(_c(dv)),
                        _p(Ks), _p(Ts), n, opnd.M, D, opnd.Do, _p(dx), _stream(x)), 'det_dx')
    return dx


def det_syrk(prec, Ks, dv, M):
    """aep_models.py:493  dB[d] = sum_n dv[n,d] kfu kfu^T."""
    lib = _lib.get()
    n, Do = dv.shape
    dB = torch.empty((Do, M, M), dtype=torch.float64, device=dv.device)
    ws = _ws(lib.gpb_det_syrk_ws_bytes(n, M, Do), dv)
    _chk(lib.gpb_det_syrk(prec, _p(Ks), _p(_c(dv)), n, M, Do, _p(dB), _p(ws), ws.numel(),
                          _stream(dv)), 'det_syrk')
    return dB


def mm_fwd(prec, mx, vx, z, ls, sf, A, B, save=True):
    """aep_models.py:183-199 / base_models.py:286-307; psi2 stays on chip.
    Returns mout, vout and the two buffers mm_bwd reuses: vacc[n,Do] = sum_ab B[d,a,b] psi2[n,a,b]
    and (save=True) psi1[n,M]."""
    lib = _lib.get()
    n, Q = mx.shape
    Do, M = A.shape
    mout = torch.empty((n, Do), dtype=torch.float64, device=mx.device)
    vout = torch.empty((n, Do), dtype=torch.float64, device=mx.device)
    vacc = torch.empty((n, Do), dtype=torch.float64, device=mx.device)
    psi1 = torch.empty((n, M), dtype=torch.float64, device=mx.device) if save else None
    ws = _ws(lib.gpb_mm_ws_bytes(n, M, Q, Do, 0), mx)
    _chk(lib.gpb_mm_fwd(prec, _p(_c(mx)), _p(_c(vx)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), _p(_c(A)),
                        _p(_c(B)), n, M, Q, Do, _p(mout), _p(vout), _p(vacc), _p(psi1), _p(ws), ws.numel(),
                        _stream(mx)), 'mm_fwd')
    return mout, vout, vacc, psi1


def mm_bwd(prec, mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1):
    """aep_models.py:238-250 + kernels.py:302-309,355-378,402-444.
    mout, vacc, psi1: outputs of mm_fwd(save=True) on the same inputs and the same B."""
    lib = _lib.get()
    n, Q = mx.shape
    Do, M = A.shape
    dev = mx.device
    f = torch.float64
    out = {
        'dA': torch.empty((Do, M), dtype=f, device=dev),
        'dB': torch.empty((Do, M, M), dtype=f, device=dev),
        'dzu': torch.empty((M, Q), dtype=f, device=dev),
        'dl': torch.empty((Q,), dtype=f, device=dev),
        'dsf2': torch.empty((1,), dtype=f, device=dev),
        'dvsum': torch.empty((1,), dtype=f, device=dev),
        'dmx': torch.empty((n, Q), dtype=f, device=dev),
        'dvx': torch.empty((n, Q), dtype=f, device=dev),
    }
    ws = _ws(lib.gpb_mm_ws_bytes(n, M, Q, Do, 1), mx)
    _chk(lib.gpb_mm_bwd(prec, _p(_c(mx)), _p(_c(vx)), _p(_c(z)), _p(_c(ls)), _p(_c(sf)), _p(_c(A)),
                        _p(_c(B)), _p(_c(dm)), _p(_c(dv)), _p(_c(mout)), _p(_c(vacc)), _p(_c(psi1)), n, M, Q, Do,
                        _p(out['dA']), _p(out['dB']), _p(out['dzu']), _p(out['dl']), _p(out['dsf2']),
                        _p(out['dvsum']), _p(out['dmx']), _p(out['dvx']), _p(ws), ws.numel(),
                        _stream(mx)), 'mm_bwd')
    return out


def fma_peak(prec, iters, device, blocks_per_sm=0):
    """Launch the FMA microbenchmark; returns the flop count (time it with CUDA events).
    blocks_per_sm (1..32) overrides the default of 8 resident 256-thread blocks per SM."""
    lib = _lib.get()
    sink = torch.zeros(32 * lib.gpb_sm_count(), dtype=torch.float64, device=device)
    flops = ctypes.c_double(float(blocks_per_sm))
    _chk(lib.gpb_fma_peak(prec, int(iters), _p(sink), ctypes.byref(flops), _stream(sink)), 'fma_peak')
    return flops.value


PROFILE_SLOTS = ('det_fwd', 'det_bwd', 'det_syrk', 'mm_pairs_fwd', 'mm_pairs_bwd', 'mm_rows_bwd',
                 'mm_cols_bwd', 'mm_psi1_fwd')


def profile_enable(on):
    _lib.get().gpb_profile_enable(1 if on else 0)


def profile_collect():
    """-> {slot name: (total ms, launches)} since the last collect (synchronises the events)."""
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_long * 8)()
    _lib.get().gpb_profile_collect(ms, cnt)
    return {PROFILE_SLOTS[i]: (ms[i], cnt[i]) for i in range(8)}


# ---- a12 / a13: elementwise latent-variable kernels (csrc/gpb_latent.cuh) --------------------
def _sel_args(sel, lo):
    if sel is None:
        return None, int(lo)
    if sel.dtype != torch.int64 or not sel.is_contiguous():
        raise RuntimeError('geepee_b200: row selection must be a contiguous int64 device tensor')
    return ctypes.c_void_p(sel.data_ptr()), 0


def lvm_x_fwd(mode, nat, x1, x2, sel, lo, n, prior1, prior2, alpha):
    """aep_models.py:840-861 (mode 0: cavity of x) / base_models.py:765-775 (mode 1: posterior of x) for the
    rows `sel` (device int64) or lo..lo+n-1 of the raw [N,Q] parameters.  -> m, v [n,Q]."""
    lib = _lib.get()
    Q = x1.shape[1]
    m = torch.empty((n, Q), dtype=torch.float64, device=x1.device)
    v = torch.empty((n, Q), dtype=torch.float64, device=x1.device)
    sp, lo = _sel_args(sel, lo)
    _chk(lib.gpb_lvm_x_fwd(int(mode), int(bool(nat)), _p(_c(x1)), _p(_c(x2)), sp, lo, int(n), Q, float(prior1),
                           float(prior2), float(alpha), _p(m), _p(v), _stream(x1)), 'lvm_x_fwd')
    return m, v


def lvm_x_bwd(mode, nat, x1, x2, sel, lo, n, prior1, prior2, alpha, s_cav, s_post, dmx, dvx):
    """aep_models.py:785-801,817-838,863-867 + base_models.py:913-929 (mode 0) / vfe_models.py:826-840,857-863
    (mode 1).  -> gx1, gx2 [N,Q] (zero outside the selection), sums[2]."""
    lib = _lib.get()
    N, Q = x1.shape
    dev = x1.device
    gx1 = torch.empty((N, Q), dtype=torch.float64, device=dev)
    gx2 = torch.empty((N, Q), dtype=torch.float64, device=dev)
    sums = torch.empty(2, dtype=torch.float64, device=dev)
    ws = _ws(lib.gpb_latent_ws_bytes(n * Q), x1)
    sp, lo = _sel_args(sel, lo)
    _chk(lib.gpb_lvm_x_bwd(int(mode), int(bool(nat)), _p(_c(x1)), _p(_c(x2)), sp, lo, int(n), int(N), Q,
                           float(prior1), float(prior2), float(alpha), float(s_cav), float(s_post), _p(_c(dmx)),
                           _p(_c(dvx)), _p(gx1), _p(gx2), _p(sums), _p(ws), ws.numel(), _stream(x1)), 'lvm_x_bwd')
    return gx1, gx2, sums


def ssm_cavity(xf1, xf2, prior1, prior2, alpha):
    """aep_models.py:1376-1387 over all T latent states -> cav_m, cav_v [T,Q]."""
    lib = _lib.get()
    T, Q = xf1.shape
    cm, cv = torch.empty_like(xf1), torch.empty_like(xf1)
    _chk(lib.gpb_ssm_cavity(_p(_c(xf1)), _p(_c(xf2)), int(T), Q, float(prior1), float(prior2), float(alpha),
                            _p(cm), _p(cv), _stream(xf1)), 'ssm_cavity')
    return cm, cv


def ssm_transition(mt, vt, mp, vp, sn, alpha, s_dyn):
    """aep_models.py:1334-1348 -> dm for the layer (= -dmt), dvt, sums[2] = {sum lz, sum dvt}."""
    lib = _lib.get()
    total = mp.numel()
    dm, dv = torch.empty_like(mp), torch.empty_like(mp)
    sums = torch.empty(2, dtype=torch.float64, device=mp.device)
    ws = _ws(lib.gpb_latent_ws_bytes(total), mp)
    _chk(lib.gpb_ssm_transition(_p(_c(mt)), _p(_c(vt)), _p(_c(mp)), _p(_c(vp)), _p(_c(sn)), int(total), float(alpha),
                                float(s_dyn), _p(dm), _p(dv), _p(sums), _p(ws), ws.numel(), _stream(mp)),
         'ssm_transition')
    return dm, dv, sums


def ssm_sources(xf1, xf2, prior1, prior2, alpha, prev, nxt, up):
    """aep_models.py:1234-1285 -> l1, l2 [T,Q].  prev / nxt / up: None or (dm, dv, first_row) with dm, dv
    [count, ld >= Q] contiguous."""
    lib = _lib.get()
    T, Q = xf1.shape
    l1, l2 = torch.empty_like(xf1), torch.empty_like(xf1)
    args = []
    for s in (prev, nxt, up):
        if s is None:
            args += [None, None, 0, 0, Q]
        else:
            dm, dv, first = s
            args += [_p(_c(dm)), _p(_c(dv)), int(first), int(dm.shape[0]), int(dm.shape[1])]
    _chk(lib.gpb_ssm_sources(_p(_c(xf1)), _p(_c(xf2)), int(T), Q, float(prior1), float(prior2), float(alpha),
                             *args, _p(l1), _p(l2), _stream(xf1)), 'ssm_sources')
    return l1, l2


def ssm_xfinal(xf1, xf2, prior1, prior2, alpha, l1, l2):
    """aep_models.py:1208-1232,1287-1315,1389-1437 -> gx1, gx2 [T,Q], sums[2] = {phi_post, phi_cav}."""
    lib = _lib.get()
    T, Q = xf1.shape
    gx1, gx2 = torch.empty_like(xf1), torch.empty_like(xf1)
    sums = torch.empty(2, dtype=torch.float64, device=xf1.device)
    ws = _ws(lib.gpb_latent_ws_bytes(T * Q), xf1)
    _chk(lib.gpb_ssm_xfinal(_p(_c(xf1)), _p(_c(xf2)), int(T), Q, float(prior1), float(prior2), float(alpha),
                            _p(_c(l1)), _p(_c(l2)), _p(gx1), _p(gx2), _p(sums), _p(ws), ws.numel(), _stream(xf1)),
         'ssm_xfinal')
    return gx1, gx2, sums
