"""oracle/geepee_oracle.py -- TEST INFRASTRUCTURE, not product code.

CPU (numpy, IEEE fp64) restatement of the reference's algorithm for the hot
path: the per-minibatch AEP / VFE energy-and-gradient evaluation of the
sparse-GP layer family (SGPR, SDGPR, SGPLVM, SGPSSM) of thangbui/geepee.
Every function cites the reference file:line it follows (paths relative to
/root/reference/geepee/).  The heavy contractions deliberately keep the
reference's un-optimised multi-operand ``np.einsum`` forms: they ARE the
reference's CPU algorithm and this module doubles as the "port" CPU baseline.

Pinning: the reference ships no golden vectors (its tests only print finite-
difference mismatches), so this module is pinned against outputs of the
reference itself: ``oracle/make_ref.py`` builds a mechanically py3-patched copy
of the reference in this container, ``tests/golden/gen_golden.py`` runs it on
seeded inputs and commits energy + every gradient as ``tests/golden/*.npz``,
and ``tests/test_oracle.py`` checks this module against those files (<=1e-9
relative; observed <=1e-12).  The reference's own identities (AEP(alpha->0) ==
VFE, psi(vx=0) == kernel, finite-difference gradients) are tested as well.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product (geepee_b200)
never does, and has no CPU fallback.
"""
import ctypes
import os

import numpy as np

JITTER = 1e-5            # config.py:11
PROP_MM = 'MM'           # config.py:13
PROP_MC = 'MC'           # config.py:15
MC_NO_SAMPLES = 5        # config.py:16

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def _clib():
    """C restatement of the weave loop (oracle/psi_oracle.c), built on demand."""
    global _CLIB
    if _CLIB is None:
        so = os.path.join(_HERE, '_ref', 'libgeepee_oracle.so')
        src = os.path.join(_HERE, 'psi_oracle.c')
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            import subprocess
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', so, src, '-lm'])
        _CLIB = ctypes.CDLL(so)
        _CLIB.geepee_oracle_psi.restype = None
        _CLIB.geepee_oracle_kernel.restype = None
    return _CLIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------
# L0: kernels.py
# --------------------------------------------------------------------------
def ard_kernel(lls2, lsf2, x, z):
    """kernels.py:10-22 compute_kernel(lls, lsf, x, z), called by the layers with
    lls = 2*ls (log squared lengthscales) and lsf = 2*sf (log signal variance).
    scipy's cdist('seuclidean', V)**2 is sum_q (x_q-z_q)^2 / V_q."""
    ls2 = np.exp(lls2)
    sf2 = np.exp(lsf2)
    x = np.atleast_2d(x)
    z = np.atleast_2d(z)
    diff = x[:, None, :] - z[None, :, :]
    r2 = np.sum(diff * diff / ls2, axis=2)
    return sf2 * np.exp(-0.5 * r2)


def psi_stats(lls2, lsf2, mx, vx, z):
    """kernels.py:181-240 compute_psi_weave: psi1[N,M], psi2[N,M,M] of the ARD-SE
    kernel under diagonal-Gaussian inputs.  Python prologue as lines 182-195, the
    inline C++ body (201-234) is run from oracle/psi_oracle.c."""
    ls2 = np.ascontiguousarray(np.exp(lls2), dtype=np.float64)
    sf2 = float(np.ravel(np.exp(lsf2))[0])
    mx = np.ascontiguousarray(mx, dtype=np.float64)
    vx = np.ascontiguousarray(vx, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    N, Q = mx.shape
    M = z.shape[0]
    ld2 = np.ascontiguousarray(0.5 * np.log(ls2 / (ls2 + 2.0 * vx)))
    ld1 = np.ascontiguousarray(0.5 * np.log(ls2 / (ls2 + vx)))
    psi1 = np.empty((N, M))
    psi2 = np.empty((N, M, M))
    _clib().geepee_oracle_psi(
        ctypes.c_long(N), ctypes.c_long(M), ctypes.c_long(Q), ctypes.c_double(sf2),
        _dp(ls2), _dp(z), _dp(mx), _dp(vx), _dp(ld1), _dp(ld2), _dp(psi1), _dp(psi2))
    return psi1, psi2


def psi_stats_numpy(lls2, lsf2, mx, vx, z):
    """Pure-numpy twin (kernels.py:317-352, the GPy-derived psi1computations /
    psi2computations); used to cross-check the C loop."""
    ls2 = np.exp(lls2)
    sf2 = np.ravel(np.exp(lsf2))[0]
    d1 = 1.0 / (vx + ls2)
    lg1 = np.log(vx / ls2 + 1.0).sum(-1)
    e1 = np.einsum('nmq,nq->nm', np.square(mx[:, None, :] - z[None]), d1)
    psi1 = sf2 * np.exp(-0.5 * (lg1[:, None] + e1))
    d2 = 1.0 / (2.0 * vx + ls2)
    lg2 = -0.5 * np.log(2.0 * vx / ls2 + 1.0).sum(-1)
    zz = -0.25 * (np.square(z[:, None, :] - z[None]) / ls2).sum(-1)
    zh = 0.5 * (z[:, None, :] + z[None])
    e2 = np.einsum('nabq,nq->nab', np.square(mx[:, None, None, :] - zh[None]), d2)
    psi2 = sf2 * sf2 * np.exp(lg2[:, None, None] + zz[None] - e2)
    return psi1, psi2


def kfu_derivs(dkfu, kfu, ls, sf2, x, z, grad_x=False):
    """kernels.py:381-399 kfucompDer: ls = lengthscale (not squared), sf2 = variance."""
    L = dkfu * kfu
    zx = z[None, :, :] - x[:, None, :]
    dvar = L.sum() / sf2
    dz = -np.einsum('nm,nmq->mq', L, zx / ls**2)
    dl = np.einsum('nm,nmq->q', L, np.square(zx) / ls**3)
    if grad_x:
        return dvar, dl, dz, np.einsum('nm,nmq->nq', L, zx / ls**2)
    return dvar, dl, dz


def psi1_derivs(dpsi1, psi1, sf2, ls, z, mu, S):
    """kernels.py:355-378 psi1compDer."""
    l2 = np.square(ls)
    L = dpsi1 * psi1
    zm = z[None, :, :] - mu[:, None, :]
    den = 1.0 / (S + l2)
    zm2d = np.square(zm) * den[:, None, :]
    dvar = L.sum() / sf2
    dmu = np.einsum('nm,nmq,nq->nq', L, zm, den)
    dS = np.einsum('nm,nmq,nq->nq', L, zm2d - 1.0, den) / 2.0
    dz = -np.einsum('nm,nmq,nq->mq', L, zm, den)
    dl = np.einsum('nm,nmq,nq->q', L, zm2d + (S / l2)[:, None, :], den * ls)
    return dvar, dl, dz, dmu, dS


def psi2_derivs(dpsi2, psi2, sf2, ls, z, mu, S):
    """kernels.py:402-444 psi2compDer (dpsi2 is [N,M,M]; symmetrised as 415-418)."""
    N, M, Q = mu.shape[0], z.shape[0], mu.shape[1]
    l2 = np.square(ls)
    den = 1.0 / (2.0 * S + l2)
    den2 = np.square(den)
    dpsi2 = 0.5 * (dpsi2 + np.swapaxes(dpsi2, 1, 2))
    L = dpsi2 * psi2
    Ls = L.reshape(N, M * M).sum(1)
    tmp = L.reshape(N * M, M).dot(z).reshape(N, M, Q)
    LZ = tmp.sum(1)
    LZ2 = L.reshape(N * M, M).dot(np.square(z)).reshape(N, M, Q).sum(1)
    LZ2p = (tmp * z[None]).sum(1)
    LZh2 = 0.5 * (LZ2 + LZ2p)
    dvar = 2.0 * Ls.sum() / sf2
    dmu = (-2.0 * den) * (mu * Ls[:, None] - LZ)
    dS = 2.0 * den2 * (np.square(mu) * Ls[:, None] - 2.0 * mu * LZ + LZh2) - den * Ls[:, None]
    L_N = L.sum(0)
    L_M = L.sum(2)
    dz = (-L_N.sum(0)[:, None] * z / l2 + L_N.dot(z) / l2
          + 2.0 * L_M.T.dot(mu * den) - L_M.T.dot(den) * z
          - (L.reshape(N, M * M).T.dot(den).reshape(M, M, Q) * z[None]).sum(1))
    dl = 2.0 * ls * ((S / l2 * den + np.square(mu * den)) * Ls[:, None]
                     + (LZ2 - LZ2p) / (2.0 * np.square(l2))
                     - (2.0 * mu * den2) * LZ + den2 * LZh2).sum(0)
    return dvar, dl, dz, dmu, dS


def psi_derivs(dpsi1, psi1, dpsi2, psi2, ls, sf2, mx, vx, z):
    """kernels.py:302-309 compute_psi_derivatives."""
    a = psi1_derivs(dpsi1, psi1, sf2, ls, z, mx, vx)
    b = psi2_derivs(dpsi2, psi2, sf2, ls, z, mx, vx)
    return tuple(p + q for p, q in zip(a, b))


def dtrace_MKzz(lls2, lsf2, z, Mm, Kzz):
    """kernels.py:447-475 d_trace_MKzz_dhypers: derivatives of tr(Mm^T Kzz) wrt
    log sf2, log ls2 and z (Mm need not be symmetric)."""
    ls2 = np.exp(lls2)
    g_sf = np.sum(Mm * Kzz)
    Ml = 0.5 * Mm * Kzz
    Xl = z / np.sqrt(ls2)
    ones = np.ones(z.shape[0])
    g_ls = ones.dot(Ml.T.dot(Xl**2)) + ones.dot(Ml.dot(Xl**2)) - 2.0 * ones.dot(Xl * Ml.dot(Xl))
    Xb = z / ls2
    g_z = 0.0
    for Mb in (-Mm.T * Kzz, -Mm * Kzz):
        g_z = g_z + Xb * Mb.sum(0)[:, None] - Mb.dot(Xb)
    return g_sf, g_ls, g_z


# --------------------------------------------------------------------------
# L1: the sparse-GP layer (base_models.py:168-658, aep_models.py:26-586,
# vfe_models.py:290-548)
# --------------------------------------------------------------------------
def _triu_pack_grad(R, dtheta1):
    """Chain rule theta_1 = R^T R with log-diagonal packing
    (aep_models.py:572-584 / base_models.py:505-514)."""
    Dout, M, _ = R.shape
    dR = np.einsum('dab,dbc->dac', R, dtheta1 + np.transpose(dtheta1, [0, 2, 1]))
    iu = np.triu_indices(M)
    di = np.diag_indices(M)
    out = np.zeros((Dout, M * (M + 1) // 2))
    for d in range(Dout):
        g = dR[d].copy()
        g[di] = g[di] * R[d][di]
        out[d] = g[iu]
    return out


class Layer(object):
    def __init__(self, N, Din, Dout, M, nat_param=True):
        self.N, self.Din, self.Dout, self.M = N, Din, Dout, M
        self.nat_param = nat_param

    # ---- base_models.py:630-658 update_hypers ----------------------------
    def set_params(self, p, suffix=''):
        M, Dout = self.M, self.Dout
        self.ls = p['ls' + suffix]
        self.sf = p['sf' + suffix]
        self.zu = p['zu' + suffix]
        iu = np.triu_indices(M)
        di = np.diag_indices(M)
        self.R = np.zeros((Dout, M, M))
        self.theta_1 = np.zeros((Dout, M, M))
        self.theta_2 = np.array(p['eta2' + suffix], dtype=np.float64)
        for d in range(Dout):
            R = np.zeros((M, M))
            R[iu] = p['eta1_R' + suffix][d]
            R[di] = np.exp(R[di])
            self.R[d] = R
            self.theta_1[d] = R.T.dot(R)
        # base_models.py:454-464 compute_kuu
        self.Kuu = ard_kernel(2 * self.ls, 2 * self.sf, self.zu, self.zu) + JITTER * np.eye(M)
        self.Kuuinv = np.linalg.inv(self.Kuu)
        self.posterior()

    # ---- base_models.py:466-488 update_posterior -------------------------
    def posterior(self):
        Ki = self.Kuuinv
        if self.nat_param:
            self.Suinv = Ki + self.theta_1
            self.Su = np.linalg.inv(self.Suinv)
            self.mu = np.einsum('dab,db->da', self.Su, self.theta_2)
        else:
            self.Su = self.theta_1
            self.Suinv = np.linalg.inv(self.Su)
            self.mu = self.theta_2
        self.Spmm = self.Su + np.einsum('da,db->dab', self.mu, self.mu)
        self.A = np.einsum('ab,db->da', Ki, self.mu)
        self.B_sto = -Ki + np.einsum('ab,dbc->dac', Ki, np.einsum('dab,bc->dac', self.Spmm, Ki))
        self.B_det = -Ki + np.einsum('ab,dbc->dac', Ki, np.einsum('dab,bc->dac', self.Su, Ki))

    # ---- aep_models.py:513-546 compute_cavity ----------------------------
    def cavity(self, alpha):
        Ki = self.Kuuinv
        beta = (self.N - alpha) * 1.0 / self.N
        if self.nat_param:
            self.Suhatinv = Ki + beta * self.theta_1
            self.Suhat = np.linalg.inv(self.Suhatinv)
            self.muhat = np.einsum('dab,db->da', self.Suhat, beta * self.theta_2)
        else:
            f1 = self.Suinv - Ki
            f2 = np.einsum('dab,db->da', self.Suinv, self.mu)
            self.Suhatinv = Ki + beta * f1
            self.Suhat = np.linalg.inv(self.Suhatinv)
            self.muhat = np.einsum('dab,db->da', self.Suhat, beta * f2)
        self.Ahat = np.einsum('ab,db->da', Ki, self.muhat)
        self.Spmmhat = self.Suhat + np.einsum('da,db->dab', self.muhat, self.muhat)
        self.Bhat_sto = -Ki + np.einsum('ab,dbc->dac', Ki, np.einsum('dab,bc->dac', self.Spmmhat, Ki))
        self.Bhat_det = -Ki + np.einsum('ab,dbc->dac', Ki, np.einsum('dab,bc->dac', self.Suhat, Ki))

    # ---- forward: aep_models.py:142-158,183-199; base_models.py:265-307 ---
    def prop_det(self, x, cav=True):
        A, B = (self.Ahat, self.Bhat_det) if cav else (self.A, self.B_det)
        kfu = ard_kernel(2 * self.ls, 2 * self.sf, x, self.zu)
        mout = np.einsum('nm,dm->nd', kfu, A)
        vout = np.exp(2 * self.sf) + np.einsum('dab,na,nb->nd', B, kfu, kfu)
        return mout, vout, kfu

    def prop_mm(self, mx, vx, cav=True):
        A, B = (self.Ahat, self.Bhat_sto) if cav else (self.A, self.B_sto)
        psi1, psi2 = psi_stats(2 * self.ls, 2 * self.sf, mx, vx, self.zu)
        mout = np.einsum('nm,dm->nd', psi1, A)
        vout = np.exp(2.0 * self.sf) + np.einsum('dab,nab->nd', B, psi2) - mout**2
        return mout, vout, psi1, psi2

    # ---- Monte-Carlo propagation: aep_models.py:160-180, base_models.py:309-332 ----
    def prop_mc(self, mx, vx, cav=True, K=None):
        """Samples x = mx + sqrt(vx) eps with eps from the GLOBAL numpy RNG, exactly as the
        reference draws it; returns 3-D (m, v) and the stacked intermediates."""
        K = MC_NO_SAMPLES if K is None else K
        n = mx.shape[0]
        eps = np.random.randn(K, n, self.Din)
        x = eps * np.sqrt(vx) + mx
        xs = x.reshape(K * n, self.Din)
        ms, vs, kfus = self.prop_det(xs, cav=cav)
        return ms.reshape(K, n, self.Dout), vs.reshape(K, n, self.Dout), (ms, vs, kfus, xs, eps)

    @staticmethod
    def reparam(dx, v, eps):
        """base_models.py:373-388 backprop_grads_reparam."""
        dx = dx.reshape(eps.shape)
        return {'mx': np.sum(dx, axis=0), 'vx': np.sum(dx * eps, axis=0) / (2 * np.sqrt(v))}

    def aep_grads_mc(self, m, v, dm, dv, kfu, x, alpha):
        """aep_models.py:307-410 backprop_grads_lvm_mc (stacked samples; natural parameters)."""
        N, Ki = self.N, self.Kuuinv
        ls, sf2 = np.exp(self.ls), np.exp(2 * self.sf)
        dm, dv = dm.reshape(m.shape), dv.reshape(v.shape)
        beta = (N - alpha) * 1.0 / N
        s_post, s_cav = N * 1.0 / alpha - 1.0, -N * 1.0 / alpha
        dkfu = np.einsum('nd,dm->nm', dm, self.Ahat) + 2 * np.einsum('nd,dab,na->nb', dv, self.Bhat_det, kfu)
        dsf2, dls, dzu, dx = kfu_derivs(dkfu, kfu, ls, sf2, x, self.zu, grad_x=True)
        kK = kfu.dot(Ki)
        SK = np.einsum('dab,nb->nda', self.Suhat, kK)
        dSinv = -np.einsum('nda,nd,ndb->dab', SK, dv, SK) - np.einsum('nda,nd,db->dab', SK, dm, self.muhat)
        dtheta1 = beta * dSinv - 0.5 * s_post * self.Spmm - 0.5 * s_cav * beta * self.Spmmhat
        dtheta2 = beta * np.einsum('nda,nd->da', SK, dm) + s_post * self.mu + s_cav * beta * self.muhat
        dA = np.einsum('nd,nm->dm', dm, kfu)
        dB = np.einsum('nd,na,nb->dab', dv, kfu, kfu)
        KS = np.einsum('ab,dbc->dac', Ki, self.Suhat)
        dKi = np.einsum('da,db->ab', dA, self.muhat) + 2 * np.einsum('dab,dac->bc', KS, dB) \
            - np.sum(dB, axis=0) + np.sum(dSinv, axis=0)
        Minner = s_post * np.sum(self.Spmm, axis=0) + s_cav * np.sum(self.Spmmhat, axis=0) - 2.0 * dKi
        M_all = 0.5 * (self.Dout * Ki + Ki.dot(Minner).dot(Ki))
        dsf, dls, dzu = self._kernel_hyper_tail(dsf2, dls, dzu, np.sum(dv), M_all)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': _triu_pack_grad(self.R, dtheta1),
                'eta2': dtheta2}, dx

    def vfe_grads_mc(self, m, v, dm, dv, kfu, x):
        """vfe_models.py:405-476 backprop_grads_lvm_mc: the deterministic-input chain rules
        (vfe_grads_det) on the stacked samples + the gradient wrt the samples."""
        dm, dv = dm.reshape(m.shape), dv.reshape(v.shape)
        g = self.vfe_grads_det(m, v, dm, dv, kfu, x)
        dkfu = np.einsum('nd,dm->nm', dm, self.A) + 2 * np.einsum('nd,dab,na->nb', dv, self.B_det, kfu)
        dx = kfu_derivs(dkfu, kfu, np.exp(self.ls), np.exp(2 * self.sf), x, self.zu, grad_x=True)[3]
        return g, dx

    # ---- log-partitions: aep_models.py:62-114 ----------------------------
    def phi(self, alpha):
        N = self.N
        phi_prior = self.Dout * 0.5 * np.linalg.slogdet(self.Kuu)[1]
        phi_post = 0.5 * np.sum(np.linalg.slogdet(self.Su)[1]) + 0.5 * np.sum(
            self.mu * np.linalg.solve(self.Su, self.mu[..., None])[..., 0])
        phi_cav = 0.5 * np.sum(np.linalg.slogdet(self.Suhat)[1]) + 0.5 * np.sum(
            self.muhat * np.linalg.solve(self.Suhat, self.muhat[..., None])[..., 0])
        return phi_prior + (N * 1.0 / alpha - 1.0) * phi_post - (N * 1.0 / alpha) * phi_cav

    # ---- vfe_models.py:309-325 compute_KL --------------------------------
    def kl(self):
        ld_prior = self.Dout * np.linalg.slogdet(self.Kuu)[1]
        ld_post = np.sum(np.linalg.slogdet(self.Su)[1])
        tr = np.sum(self.Kuuinv * self.Spmm)
        return 0.5 * (ld_prior - ld_post - self.Dout * self.M + tr)

    # ---- base_models.py:490-516 compute_posterior_grad_u -----------------
    def post_grad_u(self, dmu, dSu):
        if self.nat_param:
            dSu = dSu + np.einsum('da,db->dab', dmu, self.theta_2)
            dSuinv = -np.einsum('dab,dbc,dce->dae', self.Su, dSu, self.Su)
            dKi = np.sum(dSuinv, axis=0)
            dtheta1 = dSuinv
            deta2 = np.einsum('dab,db->da', self.Su, dmu)
        else:
            deta2, dtheta1, dKi = dmu, dSu, 0
        return _triu_pack_grad(self.R, dtheta1), deta2, dKi

    # ---- aep_models.py:548-586 compute_cav_grad_u ------------------------
    def cav_grad_u(self, dmu, dSu, alpha):
        beta = (self.N - alpha) * 1.0 / self.N
        if self.nat_param:
            dSu = dSu + np.einsum('da,db->dab', dmu, beta * self.theta_2)
            dSuinv = -np.einsum('dab,dbc,dce->dae', self.Suhat, dSu, self.Suhat)
            dKi = np.sum(dSuinv, axis=0)
            dtheta1 = beta * dSuinv
            deta2 = beta * np.einsum('dab,db->da', self.Suhat, dmu)
        else:
            f2 = np.einsum('dab,db->da', self.Suinv, self.mu)
            dSuhat = dSu + np.einsum('da,db->dab', dmu, beta * f2)
            dSuhatinv = -np.einsum('dab,dbc,dce->dae', self.Suhat, dSuhat, self.Suhat)
            dSuinv_1 = beta * dSuhatinv
            Sdm = np.einsum('dab,db->da', self.Suhat, dmu)
            dSuinv = dSuinv_1 + beta * np.einsum('da,db->dab', Sdm, self.mu)
            dtheta1 = -np.einsum('dab,dbc,dce->dae', self.Suinv, dSuinv, self.Suinv)
            deta2 = beta * np.einsum('dab,db->da', self.Suinv, Sdm)
            dKi = (1 - beta) / beta * np.sum(dSuinv_1, axis=0)
        return _triu_pack_grad(self.R, dtheta1), deta2, dKi

    def _kernel_hyper_tail(self, dsf2, dls, dzu, dv_sum, Mm):
        """aep_models.py:455-460,497-504 (same in every backprop_*): fold the direct
        kernel derivatives with d tr(Mm Kzz), Kzz = Kuu - JITTER*I."""
        ls = np.exp(self.ls)
        sf2 = np.exp(2 * self.sf)
        dls = dls * ls
        dsf = 2 * sf2 * (dsf2 + dv_sum)
        h = dtrace_MKzz(2 * self.ls, 2 * self.sf, self.zu, Mm,
                        self.Kuu - JITTER * np.eye(self.M))
        return dsf + 2 * h[0], dls + 2 * h[1], dzu + h[2]

    # ---- aep_models.py:413-511 backprop_grads_reg ------------------------
    def aep_grads_det(self, m, v, dm, dv, kfu, x, alpha):
        N = self.N
        Ki = self.Kuuinv
        scale_post = N * 1.0 / alpha - 1.0
        scale_cav = -N * 1.0 / alpha
        dkfu = np.einsum('nd,dm->nm', dm, self.Ahat) \
            + 2 * np.einsum('nd,dab,na->nb', dv, self.Bhat_det, kfu)
        dsf2, dls, dzu = kfu_derivs(dkfu, kfu, np.exp(self.ls), np.exp(2 * self.sf), x, self.zu)
        kK = np.dot(kfu, Ki)
        dmucav = np.einsum('nd,nm->dm', dm, kK)
        dSucav = np.einsum('na,nd,nb->dab', kK, dv, kK)
        Sim = np.einsum('dab,db->da', self.Suhatinv, self.muhat)
        dmucav += scale_cav * Sim
        dSucav += scale_cav * (0.5 * self.Suhatinv - 0.5 * np.einsum('da,db->dab', Sim, Sim))
        e1c, e2c, dKi_cav = self.cav_grad_u(dmucav, dSucav, alpha)
        Sim = np.einsum('dab,db->da', self.Suinv, self.mu)
        dmu = scale_post * Sim
        dSu = scale_post * (0.5 * self.Suinv - 0.5 * np.einsum('da,db->dab', Sim, Sim))
        e1p, e2p, dKi_post = self.post_grad_u(dmu, dSu)
        dKi_phi = dKi_cav + dKi_post - 0.5 * self.Dout * self.Kuu
        dAhat = np.einsum('nd,nm->dm', dm, kfu)
        dBhat = np.einsum('nd,na,nb->dab', dv, kfu, kfu)
        KiS = np.einsum('ab,dbc->dac', Ki, self.Suhat)
        dKi = np.einsum('da,db->ab', dAhat, self.muhat) \
            + 2 * np.einsum('dab,dac->bc', KiS, dBhat) - np.sum(dBhat, axis=0) + dKi_phi
        Mm = -np.dot(Ki, np.dot(dKi, Ki))
        dsf, dls, dzu = self._kernel_hyper_tail(dsf2, dls, dzu, np.sum(dv), Mm)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': e1c + e1p, 'eta2': e2c + e2p}

    # ---- aep_models.py:202-304 backprop_grads_lvm_mm ---------------------
    def aep_grads_mm(self, m, v, dm, dv, psi1, psi2, mx, vx, alpha):
        N = self.N
        Ki = self.Kuuinv
        beta = (N - alpha) * 1.0 / N
        scale_post = N * 1.0 / alpha - 1.0
        scale_cav = -N * 1.0 / alpha
        dm_all = dm - 2 * dv * m
        dAhat = np.einsum('nd,nm->dm', dm_all, psi1)
        dBhat = np.einsum('nd,nab->dab', dv, psi2)
        dpsi1 = np.einsum('nd,dm->nm', dm_all, self.Ahat)
        dpsi2 = np.einsum('nd,dab->nab', dv, self.Bhat_sto)
        dsf2, dls, dzu, dmx, dvx = psi_derivs(
            dpsi1, psi1, dpsi2, psi2, np.exp(self.ls), np.exp(2 * self.sf), mx, vx, self.zu)
        dvcav = np.einsum('ab,dbc,ce->dae', Ki, dBhat, Ki)
        dmcav = 2 * np.einsum('dab,db->da', dvcav, self.muhat) + np.einsum('ab,db->da', Ki, dAhat)
        dvcav += beta * np.einsum('da,db->dab', dmcav, self.theta_2)
        dvcavinv = -np.einsum('dab,dbc,dce->dae', self.Suhat, dvcav, self.Suhat)
        dtheta1 = beta * dvcavinv
        dtheta2 = beta * np.einsum('dab,db->da', self.Suhat, dmcav)
        KiS = np.einsum('ab,dbc->dac', Ki, self.Spmmhat)
        dKi = np.einsum('da,db->ab', dAhat, self.muhat) \
            + 2 * np.einsum('dab,dac->bc', KiS, dBhat) - np.sum(dBhat, axis=0) \
            + np.sum(dvcavinv, axis=0)
        Minner = scale_post * np.sum(self.Spmm, axis=0) \
            + scale_cav * np.sum(self.Spmmhat, axis=0) - 2.0 * dKi
        dtheta1 = -0.5 * scale_post * self.Spmm - 0.5 * scale_cav * beta * self.Spmmhat + dtheta1
        dtheta2 = scale_post * self.mu + scale_cav * beta * self.muhat + dtheta2
        deta1_R = _triu_pack_grad(self.R, dtheta1)
        M_all = 0.5 * (self.Dout * Ki + np.dot(Ki, np.dot(Minner, Ki)))
        dsf, dls, dzu = self._kernel_hyper_tail(dsf2, dls, dzu, np.sum(dv), M_all)
        return ({'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': deta1_R, 'eta2': dtheta2},
                {'mx': dmx, 'vx': dvx})

    # ---- vfe_models.py:479-548 backprop_grads_reg ------------------------
    def vfe_grads_det(self, m, v, dm, dv, kfu, x):
        Ki = self.Kuuinv
        dkfu = np.einsum('nd,dm->nm', dm, self.A) \
            + 2 * np.einsum('nd,dab,na->nb', dv, self.B_det, kfu)
        dsf2, dls, dzu = kfu_derivs(dkfu, kfu, np.exp(self.ls), np.exp(2 * self.sf), x, self.zu)
        kK = np.dot(kfu, Ki)
        dmu = np.einsum('nd,nm->dm', dm, kK) + np.einsum('ab,db->da', Ki, self.mu)
        dSu = np.einsum('na,nd,nb->dab', kK, dv, kK) + 0.5 * (Ki - self.Suinv)
        e1, e2, dKi_u = self.post_grad_u(dmu, dSu)
        dA = np.einsum('nd,nm->dm', dm, kfu)
        dB = np.einsum('nd,na,nb->dab', dv, kfu, kfu)
        KiS = np.einsum('ab,dbc->dac', Ki, self.Su)
        dKi = np.einsum('da,db->ab', dA, self.mu) + 2 * np.einsum('dab,dac->bc', KiS, dB) \
            - np.sum(dB, axis=0) + dKi_u - 0.5 * self.Dout * self.Kuu + 0.5 * np.sum(self.Spmm, axis=0)
        Mm = -np.dot(Ki, np.dot(dKi, Ki))
        dsf, dls, dzu = self._kernel_hyper_tail(dsf2, dls, dzu, np.sum(dv), Mm)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': e1, 'eta2': e2}

    # ---- vfe_models.py:328-401 backprop_grads_lvm_mm ---------------------
    def vfe_grads_mm(self, m, v, dm, dv, psi1, psi2, mx, vx):
        Ki = self.Kuuinv
        dm_all = dm - 2 * dv * m
        dpsi1 = np.einsum('nd,dm->nm', dm_all, self.A)
        dpsi2 = np.einsum('nd,dab->nab', dv, self.B_sto)
        dsf2, dls, dzu, dmx, dvx = psi_derivs(
            dpsi1, psi1, dpsi2, psi2, np.exp(self.ls), np.exp(2 * self.sf), mx, vx, self.zu)
        dA = np.einsum('nd,nm->dm', dm_all, psi1)
        dB = np.einsum('nd,nab->dab', dv, psi2)
        dSu = np.einsum('ab,dbc,ce->dae', Ki, dB, Ki)
        dmu = 2 * np.einsum('dab,db->da', dSu, self.mu) + np.einsum('ab,db->da', Ki, dA)
        dmu += np.einsum('ab,db->da', Ki, self.mu)
        dSu += 0.5 * (Ki - self.Suinv)
        e1, e2, dKi_u = self.post_grad_u(dmu, dSu)
        KiS = np.einsum('ab,dbc->dac', Ki, self.Spmm)
        dKi = np.einsum('da,db->ab', dA, self.mu) + 2 * np.einsum('dab,dac->bc', KiS, dB) \
            - np.sum(dB, axis=0) + dKi_u - 0.5 * self.Dout * self.Kuu + 0.5 * np.sum(self.Spmm, axis=0)
        Mm = -np.dot(Ki, np.dot(dKi, Ki))
        dsf, dls, dzu = self._kernel_hyper_tail(dsf2, dls, dzu, np.sum(dv), Mm)
        return ({'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': e1, 'eta2': e2},
                {'mx': dmx, 'vx': dvx})


# --------------------------------------------------------------------------
# likelihood layers (lik_layers.py)
# --------------------------------------------------------------------------
def gauss_log_Z(sn, mout, vout, y, alpha):
    """lik_layers.py:104-133 Gauss_Layer.compute_log_Z, 2-D branch.  The reference
    adds sn2/alpha to vout IN PLACE (line 122); here a new array is returned."""
    sn2 = np.exp(2.0 * sn)
    v = vout + sn2 / alpha
    D = mout.shape[1]
    logZ = np.sum(-0.5 * (np.log(2 * np.pi * v) + (y - mout)**2 / v))
    logZ += y.shape[0] * D * (0.5 * np.log(2 * np.pi * sn2 / alpha)
                              - 0.5 * alpha * np.log(2 * np.pi * sn2))
    dm = (y - mout) / v
    dv = -0.5 / v + 0.5 * (y - mout)**2 / v**2
    return logZ, dm, dv


def gauss_log_Z_mc(sn, mout, vout, y, alpha):
    """lik_layers.py:134-150: 3-D branch (samples on axis 0), log-mean-exp over the samples."""
    sn2 = np.exp(2.0 * sn)
    vout = vout + sn2 / alpha
    lz = -0.5 * (np.log(2 * np.pi * vout) + (y - mout)**2 / vout)
    lz = lz + (0.5 * np.log(2 * np.pi * sn2 / alpha) - 0.5 * alpha * np.log(2 * np.pi * sn2))
    lmax = np.max(lz, axis=0)
    ex = np.exp(lz - lmax)
    se = np.sum(ex, axis=0)
    logZ = np.sum(lmax + np.log(se) - np.log(mout.shape[0]))
    w = ex / se
    return logZ, w * (y - mout) / vout, w * (-0.5 / vout + 0.5 * (y - mout)**2 / vout**2)


def gauss_dsn_mc(sn, mout, dv, alpha, scale):
    """lik_layers.py:154-181 backprop_grads, 3-D branch (dim_prod = batch x D)."""
    sn2 = np.exp(2.0 * sn)
    return scale * (np.sum(dv) * 2 * sn2 / alpha + mout.shape[1] * mout.shape[2] * (1 - alpha))


def gauss_dsn(sn, mout, dv, alpha, scale):
    """lik_layers.py:154-181 Gauss_Layer.backprop_grads."""
    sn2 = np.exp(2.0 * sn)
    return scale * (np.sum(dv) * 2 * sn2 / alpha + mout.shape[0] * mout.shape[1] * (1 - alpha))


def gauss_log_lik_exp(sn, mout, vout, y):
    """lik_layers.py:183-199 compute_log_lik_exp, 2-D branch."""
    sn2 = np.exp(2.0 * sn)
    e = -0.5 * np.log(2 * np.pi * sn2) - 0.5 / sn2 * (y**2 - 2 * y * mout + mout**2 + vout)
    return np.sum(e), (y - mout) / sn2, -0.5 / sn2 * np.ones_like(vout)


def gauss_dsn_log_lik_exp(sn, m, v, y, scale):
    """lik_layers.py:217-226 backprop_grads_log_lik_exp, 2-D branch."""
    sn2 = np.exp(2.0 * sn)
    return scale * np.sum(-1 + (y**2 - 2 * y * m + m**2 + v) / sn2)


GH_DEGREE = 10   # config.py:12


def probit_log_Z(mout, vout, y, alpha):
    """lik_layers.py:303-362 Probit_Layer.compute_log_Z, 2-D branch (y in {-1,+1})."""
    from scipy import special
    if alpha == 1.0:
        t = y * mout / np.sqrt(1 + vout)
        Z = 0.5 * (1 + special.erf(t / np.sqrt(2)))
        eps = 1e-16
        logZ = np.sum(np.log(Z + eps))
        dlogZ_dt = 1 / (Z + eps) / np.sqrt(2 * np.pi) * np.exp(-t**2.0 / 2)
        return logZ, dlogZ_dt * y / np.sqrt(1 + vout), dlogZ_dt * (-0.5 * y * mout / (1 + vout)**1.5)
    gh_x, gh_w = np.polynomial.hermite.hermgauss(GH_DEGREE)
    gh_x, gh_w = gh_x[:, None, None], gh_w[:, None, None]
    ts = gh_x * np.sqrt(2 * vout[None]) + mout[None]
    eps = 1e-8
    pdfs = 0.5 * (1 + special.erf(y * ts / np.sqrt(2))) + eps
    Zt = np.sum(pdfs**alpha * gh_w, axis=0) / np.sqrt(np.pi)
    logZ = np.sum(np.log(Zt))
    a = pdfs**(alpha - 1.0) * np.exp(-ts**2 / 2)
    dZdm = np.sum(gh_w * a, axis=0) * y * alpha / np.pi / np.sqrt(2)
    dZdv = np.sum(gh_w * (a * gh_x), axis=0) * y * alpha / np.pi / np.sqrt(2) / np.sqrt(2 * vout)
    return logZ, dZdm / Zt + eps, dZdv / Zt + eps


def probit_log_Z_mc(mout, vout, y, alpha):
    """lik_layers.py:364-409 Probit_Layer.compute_log_Z, 3-D branch: log-mean-exp over the samples on
    axis 0 (note eps = 1e-16 here, 1e-8 in the 2-D quadrature branch, and the `+ eps` on dm, dv)."""
    from scipy import special
    eps = 1e-16
    if alpha == 1.0:
        t = y * mout / np.sqrt(1 + vout)
        Z = 0.5 * (1 + special.erf(t / np.sqrt(2)))
        lt = np.log(Z + eps)
    else:
        gh_x, gh_w = np.polynomial.hermite.hermgauss(GH_DEGREE)
        gh_x, gh_w = gh_x[:, None, None, None], gh_w[:, None, None, None]
        ts = gh_x * np.sqrt(2 * vout) + mout
        pdfs = 0.5 * (1 + special.erf(y * ts / np.sqrt(2))) + eps
        Zt = np.sum(pdfs**alpha * gh_w, axis=0) / np.sqrt(np.pi)
        lt = np.log(Zt)
    lmax = np.max(lt, axis=0)
    ex = np.exp(lt - lmax)
    se = np.sum(ex, axis=0)
    logZ = np.sum(lmax + np.log(se) - np.log(mout.shape[0]))
    w = ex / se
    if alpha == 1.0:
        dt = 1 / (Z + eps) / np.sqrt(2 * np.pi) * np.exp(-t**2.0 / 2)
        return logZ, w * dt * y / np.sqrt(1 + vout), w * dt * (-0.5 * y * mout / (1 + vout)**1.5)
    a = pdfs**(alpha - 1.0) * np.exp(-ts**2 / 2)
    dZdm = np.sum(gh_w * a, axis=0) * y * alpha / np.pi / np.sqrt(2)
    dZdv = np.sum(gh_w * (a * gh_x), axis=0) * y * alpha / np.pi / np.sqrt(2) / np.sqrt(2 * vout)
    return logZ, w * dZdm / Zt + eps, w * dZdv / Zt + eps


def probit_log_lik_exp(m, v, y):
    """lik_layers.py:418-436 Probit_Layer.compute_log_lik_exp, 2-D branch."""
    from scipy.stats import norm
    gh_x, gh_w = np.polynomial.hermite.hermgauss(GH_DEGREE)
    gh_x, gh_w = gh_x[:, None, None], gh_w[:, None, None] / np.sqrt(np.pi)
    ts = gh_x * np.sqrt(2 * v[None]) + m[None]
    loglik = np.sum(gh_w * norm.logcdf(ts * y))
    grad_cdfs = y * gh_w * norm.pdf(ts * y) / norm.cdf(ts * y)
    return loglik, np.sum(grad_cdfs, axis=0), np.sum(grad_cdfs * 0.5 * gh_x * np.sqrt(2 / v[None]), axis=0)


def lik_log_Z(lik, params, m, v, y, alpha, scale, g, sn_key='sn'):
    """Tilted log-partition of the output likelihood + its hyper gradient into g."""
    if lik == 'Probit':
        return probit_log_Z(m, v, y, alpha)
    sn = params[sn_key]
    logZ, dm, dv = gauss_log_Z(sn, m, v, y, alpha)
    g[sn_key] = gauss_dsn(sn, m, dv, alpha, scale)
    return logZ, dm, dv


def lik_log_lik_exp(lik, params, m, v, y, scale, g, sn_key='sn'):
    if lik == 'Probit':
        return probit_log_lik_exp(m, v, y)
    sn = params[sn_key]
    ll, dm, dv = gauss_log_lik_exp(sn, m, v, y)
    g[sn_key] = gauss_dsn_log_lik_exp(sn, m, v, y, scale)
    return ll, dm, dv


class GaussEmis(object):
    """lik_layers.py:474-676 Gauss_Emis: y ~ N(C x, diag(R))."""

    def __init__(self, y, Dout, Din):
        self.y, self.N, self.Dout, self.Din = y, y.shape[0], Dout, Din

    def set_params(self, p, suffix=''):
        self.C = p['C' + suffix]
        self.R = np.exp(2 * p['R' + suffix])

    def tilted(self, mx, vx, alpha, scale, idxs):
        """lik_layers.py:573-627 compute_emission_tilted."""
        C, R, Dout = self.C, self.R, self.Dout
        Nb = mx.shape[0]
        CVC = np.einsum('da,na,ab->ndb', C, vx, C.T)
        Vy = np.diag(R / alpha) + CVC
        Yd = self.y[idxs] - np.einsum('da,na->nd', C, mx)
        VinvY = np.linalg.solve(Vy, Yd[..., None])[..., 0]
        quad = -0.5 * np.sum(Yd * VinvY)
        ICVCR = np.eye(Dout)[None] + alpha * CVC / R
        logZ = (-Nb * Dout * 0.5 * alpha * np.log(2 * np.pi) - 0.5 * Nb * alpha * np.sum(np.log(R))
                - 0.5 * np.sum(np.linalg.slogdet(ICVCR)[1]) + quad)
        Vyinv = np.linalg.inv(Vy)
        dR = (-0.5 * np.sum(np.diagonal(Vyinv, axis1=1, axis2=2), axis=0)
              + 0.5 * np.sum(VinvY**2, axis=0)) / alpha
        dR += 0.5 * Nb * (1 - alpha) / R
        dR *= 2 * R
        dSig = -0.5 * Vyinv + 0.5 * np.einsum('na,nb->nab', VinvY, VinvY)
        dC = np.einsum('na,nb->ab', VinvY, mx) + 2 * np.einsum('nc,bc,nab->ac', vx, C, dSig)
        dmx = np.einsum('na,ab->nb', VinvY, C)
        dvx = np.einsum('nab,da,db->nd', dSig, C.T, C.T)
        return logZ * scale, {'mx': dmx * scale, 'vx': dvx * scale}, {'C': dC * scale, 'R': dR * scale}

    def log_lik_exp(self, mx, vx, scale, idxs):
        """lik_layers.py:629-676 compute_emission_log_lik_exp."""
        C, R, Dout = self.C, self.R, self.Dout
        Nb = mx.shape[0]
        yb = self.y[idxs]
        Cm = np.einsum('ab,nb->na', C, mx)
        CRC = C.T.dot(np.diag(1 / R)).dot(C)
        sv = np.sum(vx, axis=0)
        logZ = (-0.5 * Nb * Dout * np.log(2 * np.pi) - 0.5 * Nb * np.sum(np.log(R))
                - 0.5 * np.sum(np.sum((yb - Cm)**2, axis=0) / R) - 0.5 * np.sum(sv * np.diag(CRC)))
        dR = (-0.5 * Nb / R + 0.5 * np.sum((yb - Cm)**2, axis=0) / R**2
              + 0.5 * np.diag(C.dot(np.diag(sv)).dot(C.T)) / R**2) * 2 * R
        dC = np.diag(1 / R).dot(np.einsum('na,nb->ab', yb - Cm, mx)) \
            - np.diag(1 / R).dot(C.dot(np.diag(sv)))
        dmx = np.einsum('ba,na->nb', C.T.dot(np.diag(1 / R)), yb - Cm)
        dvx = np.tile(-0.5 * np.diag(CRC)[None, :], [Nb, 1])
        return logZ * scale, {'mx': dmx * scale, 'vx': dvx * scale}, {'C': dC * scale, 'R': dR * scale}


# --------------------------------------------------------------------------
# L2: models.  objective(params, mb_size, alpha) -> (energy, grads)
# --------------------------------------------------------------------------
def _pick_rows(N, mb_size):
    """aep_models.py:624-630: full batch, or numpy global-RNG subset."""
    if mb_size >= N:
        return None
    return np.random.choice(N, mb_size, replace=False)


class AepSGPR(object):
    """aep_models.py:589-667."""

    def __init__(self, x, y, M, nat_param=True, lik='Gaussian'):
        self.x, self.y, self.lik = x, y, lik
        self.N, self.Din, self.Dout, self.M = y.shape[0], x.shape[1], y.shape[1], M
        self.layer = Layer(self.N, self.Din, self.Dout, M, nat_param)
        self.fixed_params = []

    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N = self.N
        idx = _pick_rows(N, mb_size)
        xb, yb = (self.x, self.y) if idx is None else (self.x[idx], self.y[idx])
        scale = -N * 1.0 / yb.shape[0] / alpha
        L = self.layer
        L.set_params(params)
        sn = params.get('sn')
        L.cavity(alpha)
        m, v, kfu = L.prop_det(xb)
        gl = {}
        logZ, dm, dv = lik_log_Z(self.lik, params, m, v, yb, alpha, scale, gl)
        g = L.aep_grads_det(m, v, scale * dm, scale * dv, kfu, xb, alpha)
        g.update(gl)
        energy = scale * logZ + L.phi(alpha)
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


class VfeSGPR(object):
    """vfe_models.py:551-632."""

    def __init__(self, x, y, M, nat_param=True, lik='Gaussian'):
        self.x, self.y, self.lik = x, y, lik
        self.N, self.Din, self.Dout, self.M = y.shape[0], x.shape[1], y.shape[1], M
        self.layer = Layer(self.N, self.Din, self.Dout, M, nat_param)
        self.fixed_params = []

    def objective_function(self, params, mb_size, alpha='not_used', prop_mode='not_used'):
        N = self.N
        idx = _pick_rows(N, mb_size)
        xb, yb = (self.x, self.y) if idx is None else (self.x[idx], self.y[idx])
        scale = -N * 1.0 / yb.shape[0]
        L = self.layer
        L.set_params(params)
        sn = params.get('sn')
        m, v, kfu = L.prop_det(xb, cav=False)
        gl = {}
        ll, dm, dv = lik_log_lik_exp(self.lik, params, m, v, yb, scale, gl)
        g = L.vfe_grads_det(m, v, scale * dm, scale * dv, kfu, xb)
        g.update(gl)
        energy = scale * ll + L.kl()
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


class AepSDGPR(object):
    """aep_models.py:870-988 (layers always nat_param: line 893)."""

    def __init__(self, x, y, Ms, hidden_sizes, lik='Gaussian'):
        self.x, self.y, self.lik = x, y, lik
        self.N, self.Din, self.Dout = y.shape[0], x.shape[1], y.shape[1]
        self.size = [self.Din] + list(hidden_sizes) + [self.Dout]
        self.L = len(self.size) - 1
        self.Ms = list(Ms) if isinstance(Ms, (list, tuple)) else [Ms] * self.L
        self.layers = [Layer(self.N, self.size[i], self.size[i + 1], self.Ms[i])
                       for i in range(self.L)]
        self.fixed_params = []

    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N = self.N
        idx = _pick_rows(N, mb_size)
        xb, yb = (self.x, self.y) if idx is None else (self.x[idx], self.y[idx])
        scale = -N * 1.0 / yb.shape[0] / alpha
        for i, L in enumerate(self.layers):
            L.set_params(params, '_%d' % i)
            L.cavity(alpha)
        sn = params.get('sn')
        ms, vs, p1, p2 = [], [], [], []
        for i, L in enumerate(self.layers):
            if i == 0:
                m, v, k = L.prop_det(xb)
                q = None
            else:
                m, v, k, q = L.prop_mm(ms[-1], vs[-1])
            ms.append(m), vs.append(v), p1.append(k), p2.append(q)
        gl = {}
        logZ, dm, dv = lik_log_Z(self.lik, params, ms[-1], vs[-1], yb, alpha, scale, gl)
        dmi, dvi = scale * dm, scale * dv
        g = {}
        for i in range(self.L - 1, -1, -1):
            L = self.layers[i]
            if i == 0:
                gh = L.aep_grads_det(ms[0], vs[0], dmi, dvi, p1[0], xb, alpha)
            else:
                gh, gi = L.aep_grads_mm(ms[i], vs[i], dmi, dvi, p1[i], p2[i],
                                        ms[i - 1], vs[i - 1], alpha)
                dmi, dvi = gi['mx'], gi['vx']
            for k, val in gh.items():
                g[k + '_%d' % i] = val
        g.update(gl)
        energy = scale * logZ + sum(L.phi(alpha) for L in self.layers)
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


class AepSDGPR_H(AepSDGPR):
    """aep_models.py:1440-1864: deep GP with a Gaussian factor per training row and hidden
    unit (natural parameters h_factor_1/2[N, size[i+1]] per hidden layer, tied twice: posterior =
    2 x factor, cavity = (2 - alpha) x factor).  The hidden variables decouple the layers: layer i
    maps the cavity of hidden layer i-1 to a Gaussian that is matched to the cavity of hidden
    layer i (compute_transition_tilted, 1710-1745).  Full batch only: compute_grads_hidden
    (1621-1653) combines [N, D] factors with batch-sized gradients."""

    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N, Ln = self.N, self.L
        assert mb_size >= N, 'SDGPR_H: the reference only supports full batches'
        xb, yb = self.x, self.y
        scale = -N * 1.0 / N / alpha
        for i, L in enumerate(self.layers):
            L.set_params(params, '_%d' % i)
            L.cavity(alpha)
        snh = np.asarray(params['sn_hidden'], dtype=np.float64)
        h1 = [np.asarray(params['h_factor_1_%d' % i], dtype=np.float64) for i in range(Ln - 1)]
        h2 = [np.exp(2.0 * np.asarray(params['h_factor_2_%d' % i], dtype=np.float64)) for i in range(Ln - 1)]
        c1 = [a * (2.0 - alpha) for a in h1]            # compute_cavity_h, 1747-1767
        c2 = [a * (2.0 - alpha) for a in h2]
        cm = [a / b for a, b in zip(c1, c2)]
        cv = [1.0 / b for b in c2]
        g, dmc, dvc = {}, [], []
        dsn = np.zeros_like(snh)
        logZ = 0.0
        for i in range(Ln - 1):
            L = self.layers[i]
            if i == 0:
                mp, vp, kfu = L.prop_det(xb)
            else:
                mp, vp, psi1, psi2 = L.prop_mm(cm[i - 1], cv[i - 1])
            # compute_transition_tilted, 1710-1745
            sn2 = np.exp(2.0 * snh[i])
            vsum = cv[i] + vp + sn2 / alpha
            md = cm[i] - mp
            lz = np.sum(-0.5 * md**2 / vsum - 0.5 * np.log(2 * np.pi * vsum)
                        + 0.5 * (1 - alpha) * np.log(2 * np.pi * sn2) - 0.5 * np.log(alpha))
            dvt = -0.5 / vsum + 0.5 * md**2 / vsum**2
            dmt = -md / vsum
            dsn[i] = scale * (np.sum(dvt) * 2 * sn2 / alpha + mp.shape[0] * self.size[i + 1] * (1 - alpha))
            logZ += scale * lz
            if i == 0:
                gh = L.aep_grads_det(mp, vp, scale * (-dmt), scale * dvt, kfu, xb, alpha)
            else:
                gh, gi = L.aep_grads_mm(mp, vp, scale * (-dmt), scale * dvt, psi1, psi2,
                                        cm[i - 1], cv[i - 1], alpha)
                dmc[i - 1] = dmc[i - 1] + gi['mx']
                dvc[i - 1] = dvc[i - 1] + gi['vx']
            for k, val in gh.items():
                g[k + '_%d' % i] = val
            dmc.append(scale * dmt)
            dvc.append(scale * dvt)
        i = Ln - 1
        L = self.layers[i]
        mp, vp, psi1, psi2 = L.prop_mm(cm[i - 1], cv[i - 1])
        gl = {}
        lzl, dm, dv = lik_log_Z(self.lik, params, mp, vp, yb, alpha, scale, gl)
        gh, gi = L.aep_grads_mm(mp, vp, scale * dm, scale * dv, psi1, psi2, cm[i - 1], cv[i - 1], alpha)
        logZ += scale * lzl
        dmc[i - 1] = dmc[i - 1] + gi['mx']
        dvc[i - 1] = dvc[i - 1] + gi['vx']
        for k, val in gh.items():
            g[k + '_%d' % i] = val
        g.update(gl)
        g['sn_hidden'] = dsn
        # compute_grads_hidden 1621-1653, compute_phi_{cavity,posterior}_h 1655-1690
        s_post, s_cav = -(1.0 - 1.0 / alpha), -1.0 / alpha
        phi_h = 0.0
        for i in range(Ln - 1):
            p1, p2 = 2.0 * h1[i], 2.0 * h2[i]
            d1 = (2 - alpha) * (dmc[i] / c2[i]) + s_cav * (2 - alpha) * (c1[i] / c2[i]) \
                + s_post * 2 * (p1 / p2)
            d2 = (2 - alpha) * (-dmc[i] * c1[i] / c2[i]**2 - dvc[i] / c2[i]**2) \
                + s_cav * (2 - alpha) * (-0.5 * c1[i]**2 / c2[i]**2 - 0.5 / c2[i]) \
                + s_post * (-p1**2 / p2**2 - 1 / p2)
            g['h_factor_1_%d' % i] = d1
            g['h_factor_2_%d' % i] = 2 * d2 * h2[i]
            phi_h += s_cav * np.sum(0.5 * (c1[i]**2 / c2[i] - np.log(c2[i])))
            phi_h += s_post * np.sum(0.5 * (p1**2 / p2 - np.log(p2)))
        energy = logZ + sum(L.phi(alpha) for L in self.layers) + phi_h
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


def _phi_x(mx, vx):
    """aep_models.py:863-867 compute_phi_x."""
    return (np.sum(0.5 * (mx**2 / vx + np.log(vx))), mx / vx,
            0.5 * (-mx**2 / vx**2 + 1 / vx))


class AepSGPLVM(object):
    """aep_models.py:670-867 + base_models.py:661-929 (nat_param=True only: the
    reference's AEP moment-matched tail has no valid nat_param=False variant)."""

    def __init__(self, y, Q, M, prior_mean=0, prior_var=1, lik='Gaussian'):
        self.y, self.lik = y, lik
        self.N, self.Dout, self.Din, self.M = y.shape[0], y.shape[1], Q, M
        self.layer = Layer(self.N, Q, self.Dout, M)
        self.prior_mean, self.prior_var = prior_mean, prior_var
        self.prior_x1, self.prior_x2 = prior_mean / prior_var, 1.0 / prior_var
        self.fixed_params = []

    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N = self.N
        if mb_size == N:
            idx = np.arange(N)
        else:
            idx = np.random.choice(N, mb_size, replace=False)
        yb = self.y[idx]
        nb = yb.shape[0]
        scale = -N * 1.0 / nb / alpha
        L = self.layer
        L.set_params(params)
        sn = params.get('sn')
        f1 = params['x1']
        f2 = np.exp(2 * params['x2'])                       # base_models.py:903-904
        post1, post2 = self.prior_x1 + f1, self.prior_x2 + f2
        L.cavity(alpha)
        c1 = self.prior_x1 + (1.0 - alpha) * f1[idx]       # aep_models.py:840-861
        c2 = self.prior_x2 + (1.0 - alpha) * f2[idx]
        mcav, vcav = c1 / c2, 1.0 / c2
        mpost, vpost = post1[idx] / post2[idx], 1.0 / post2[idx]
        gl = {}
        if prop_mode == PROP_MC:                            # aep_models.py:745-761
            m, v, (ms, vs, kfus, xs, eps) = L.prop_mc(mcav, vcav)
            if self.lik == 'Probit':
                logZ, dm, dv = probit_log_Z_mc(m, v, yb, alpha)
            else:
                logZ, dm, dv = gauss_log_Z_mc(sn, m, v, yb, alpha)
                gl['sn'] = gauss_dsn_mc(sn, m, dv, alpha, scale)
            g, dx = L.aep_grads_mc(ms, vs, scale * dm, scale * dv, kfus, xs, alpha)
            gin = L.reparam(dx, vcav, eps)
        else:
            m, v, psi1, psi2 = L.prop_mm(mcav, vcav)
            logZ, dm, dv = lik_log_Z(self.lik, params, m, v, yb, alpha, scale, gl)
            g, gin = L.aep_grads_mm(m, v, scale * dm, scale * dv, psi1, psi2, mcav, vcav, alpha)
        g.update(gl)
        # aep_models.py:785-801
        phi_prior, _, _ = _phi_x(self.prior_mean, self.prior_var)
        phi_prior *= N * self.Din
        phi_cav, dmc, dvc = _phi_x(mcav, vcav)
        phi_post, dmp, dvp = _phi_x(mpost, vpost)
        s_cav = -N * 1.0 / nb / alpha
        s_post = -N * 1.0 / nb * (1.0 - 1.0 / alpha)
        x_contrib = phi_prior + s_cav * phi_cav + s_post * phi_post
        dmc = s_cav * dmc + gin['mx']
        dvc = s_cav * dvc + gin['vx']
        dmp, dvp = s_post * dmp, s_post * dvp
        # aep_models.py:817-838 compute_cav_grad_x (nat)
        t1, t2 = mcav / vcav, 1 / vcav
        d1c = (1.0 - alpha) * dmc / t2
        d2c = (1.0 - alpha) * (-dmc * t1 / t2**2 - dvc / t2**2) * 2 * f2[idx]
        # base_models.py:913-929 compute_posterior_grad_x (nat)
        p1, p2 = post1[idx], post2[idx]
        d1p = dmp / p2
        d2p = (-dmp * p1 / p2**2 - dvp / p2**2) * 2 * f2[idx]
        g['x1'] = np.zeros_like(f1)
        g['x2'] = np.zeros_like(f2)
        g['x1'][idx] = d1c + d1p
        g['x2'][idx] = d2c + d2p
        energy = scale * logZ + x_contrib + L.phi(alpha)
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy, g                                    # NOT divided by N (line 815)


class VfeSGPLVM(object):
    """vfe_models.py:722-863."""

    def __init__(self, y, Q, M, prior_mean=0, prior_var=1, nat_param=True, lik='Gaussian'):
        self.y, self.lik = y, lik
        self.N, self.Dout, self.Din, self.M = y.shape[0], y.shape[1], Q, M
        self.nat_param = nat_param
        self.layer = Layer(self.N, Q, self.Dout, M, nat_param)
        self.prior_mean, self.prior_var = prior_mean, prior_var
        self.prior_x1, self.prior_x2 = prior_mean / prior_var, 1.0 / prior_var
        self.fixed_params = []

    def objective_function(self, params, mb_size, alpha='not_used', prop_mode=PROP_MM):
        N = self.N
        if mb_size == N:
            idx = np.arange(N)
        else:
            idx = np.random.choice(N, mb_size, replace=False)
        yb = self.y[idx]
        nb = yb.shape[0]
        scale = -N * 1.0 / nb
        L = self.layer
        L.set_params(params)
        sn = params.get('sn')
        f1 = params['x1']
        f2 = np.exp(2 * params['x2'])
        if self.nat_param:
            post1, post2 = self.prior_x1 + f1, self.prior_x2 + f2
        else:
            post1, post2 = f1 / f2, 1.0 / f2
        mx, vx = post1[idx] / post2[idx], 1.0 / post2[idx]
        gl = {}
        if prop_mode == PROP_MC:                            # vfe_models.py:793-808
            m, v, (ms, vs, kfus, xs, eps) = L.prop_mc(mx, vx, cav=False)
            K = m.shape[0]
            if self.lik == 'Probit':                        # lik_layers.py:437-455: sample average
                ll, dm, dv = probit_log_lik_exp(ms, vs, np.tile(yb, (K, 1)))
                ll, dm, dv = ll / K, dm.reshape(m.shape) / K, dv.reshape(v.shape) / K
            else:
                sn2 = np.exp(2.0 * sn)                      # lik_layers.py:209-216, 229-234
                ll = np.sum(np.mean(-0.5 * np.log(2 * np.pi * sn2) - 0.5 * ((yb - m)**2 + v) / sn2, axis=0))
                dm, dv = (yb - m) / sn2 / K, -0.5 / sn2 * np.ones_like(v) / K
                gl['sn'] = scale * np.sum(-1 + ((yb - m)**2 + v) / sn2) / K
            g, dx = L.vfe_grads_mc(ms, vs, scale * dm, scale * dv, kfus, xs)
            gin = L.reparam(dx, vx, eps)
        else:
            m, v, psi1, psi2 = L.prop_mm(mx, vx, cav=False)
            ll, dm, dv = lik_log_lik_exp(self.lik, params, m, v, yb, scale, gl)
            g, gin = L.vfe_grads_mm(m, v, scale * dm, scale * dv, psi1, psi2, mx, vx)
        g.update(gl)
        m0, v0 = self.prior_mean, self.prior_var            # vfe_models.py:857-863
        klx = np.sum(0.5 * (np.log(v0) - np.log(vx) + (vx + (mx - m0)**2) / v0 - 1))
        sx = N * 1.0 / nb
        dmx = gin['mx'] + sx * (mx - m0) / v0
        dvx = gin['vx'] + sx * (-0.5 / vx + 0.5 / v0)
        g['x1'] = np.zeros_like(f1)
        g['x2'] = np.zeros_like(f2)
        if self.nat_param:
            p1, p2 = post1[idx], post2[idx]
            g['x1'][idx] = dmx / p2
            g['x2'][idx] = (-dmx * p1 / p2**2 - dvx / p2**2) * 2 * f2[idx]
        else:
            g['x1'][idx] = dmx
            g['x2'][idx] = dvx * 2 * f2[idx]
        energy = scale * ll + sx * klx + L.kl()
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


class _SSMBase(object):
    """Shared pieces of base_models.py:1339-1752 (Base_SGPSSM)."""

    def __init__(self, y, Q, M, prior_mean=0, prior_var=1, x_control=None,
                 gp_emi=False, control_to_emi=True, nat_param=True):
        self.y = y
        self.N, self.Dout, self.Din, self.M = y.shape[0], y.shape[1], Q, M
        self.x_control = x_control
        self.Dcon_dyn = 0 if x_control is None else x_control.shape[1]
        self.Dcon_emi = self.Dcon_dyn if (x_control is not None and control_to_emi) else 0
        self.gp_emi = gp_emi
        self.nat_param = nat_param
        self.x_prior_1, self.x_prior_2 = prior_mean / prior_var, 1.0 / prior_var
        self.dyn = Layer(self.N - 1, Q + self.Dcon_dyn, Q, M, nat_param)
        if gp_emi:
            self.emi = Layer(self.N, Q + self.Dcon_emi, self.Dout, M, nat_param)
        else:
            self.emi = GaussEmis(y, self.Dout, Q + self.Dcon_emi)
        self.fixed_params = []

    def _set(self, params):
        """base_models.py:1711-1728 update_hypers."""
        self.dyn.set_params(params, '_dynamic')
        self.emi.set_params(params, '_emission')
        self.sn = params['sn']
        self.f1 = params['x_factor_1']
        self.f2 = np.exp(2 * params['x_factor_2'])
        if self.nat_param:
            self.post1, self.post2 = 3 * self.f1, 3 * self.f2
            self.post1[[0, -1]] = 2 * self.f1[[0, -1]]
            self.post2[[0, -1]] = 2 * self.f2[[0, -1]]
            self.post1[0] += self.x_prior_1
            self.post2[0] += self.x_prior_2
        else:
            self.post1, self.post2 = self.f1 / self.f2, 1.0 / self.f2

    def _window(self, mb_size):
        """aep_models.py:1045-1057: full series or one random contiguous window."""
        N = self.N
        if mb_size >= N:
            return np.arange(0, N - 1), np.arange(0, N)
        s = np.random.randint(0, N - mb_size)
        return np.arange(s, s + mb_size - 1), np.arange(s, s + mb_size)

    def _with_control(self, m, v, idxs, Dcon):
        if Dcon > 0:
            return (np.hstack((m, self.x_control[idxs])),
                    np.hstack((v, np.zeros((m.shape[0], Dcon)))))
        return m, v


class AepSGPSSM(_SSMBase):
    """aep_models.py:991-1437."""

    def __init__(self, y, Q, M, prior_mean=0, prior_var=1, x_control=None,
                 gp_emi=False, control_to_emi=True):
        super(AepSGPSSM, self).__init__(y, Q, M, prior_mean, prior_var, x_control,
                                        gp_emi, control_to_emi, True)

    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N, Q = self.N, self.Din
        dyn_idx, emi_idx = self._window(mb_size)
        yb = self.y[emi_idx]
        s_dyn = -(N - 1) * 1.0 / dyn_idx.shape[0] / alpha
        s_emi = -N * 1.0 / emi_idx.shape[0] / alpha
        self._set(params)
        self.dyn.cavity(alpha)
        if self.gp_emi:
            self.emi.cavity(alpha)
        cav1 = self.post1 - alpha * self.f1                  # aep_models.py:1376-1387
        cav2 = self.post2 - alpha * self.f2
        cav_m, cav_v = cav1 / (cav2 + 1e-16), 1.0 / (cav2 + 1e-16)
        mt, vt = cav_m[dyn_idx + 1], cav_v[dyn_idx + 1]
        mtm1, vtm1 = self._with_control(cav_m[dyn_idx], cav_v[dyn_idx], dyn_idx, self.Dcon_dyn)
        mup, vup = self._with_control(cav_m[emi_idx], cav_v[emi_idx], emi_idx, self.Dcon_emi)
        # transition factors (aep_models.py:1092-1098 / 1114-1126, 1317-1374)
        mc = prop_mode == PROP_MC
        sn2 = np.exp(2 * self.sn)
        if mc:
            mp, vp, (ms, vs, kfus, xs, eps) = self.dyn.prop_mc(mtm1, vtm1)
        else:
            mp, vp, psi1, psi2 = self.dyn.prop_mm(mtm1, vtm1)
        vsum = vt + vp + sn2 / alpha
        md = mt - mp
        lz = -0.5 * md**2 / vsum - 0.5 * np.log(1 + alpha * (vt + vp) / sn2) \
            - 0.5 * alpha * np.log(2 * np.pi * sn2)
        if mc:                                               # 3-D branch, 1349-1369
            lmax = np.max(lz, axis=0)
            ex = np.exp(lz - lmax)
            se = np.sum(ex, axis=0)
            logZ_dyn = s_dyn * np.sum(lmax + np.log(se) - np.log(mp.shape[0]))
            w = s_dyn * ex / se
            dmp = w * md / vsum
            dvp = w * (-0.5 / vsum + 0.5 * md**2 / vsum**2)
            dmt, dvt = -np.sum(dmp, axis=0), np.sum(dvp, axis=0)
            dsn = np.sum(dvp) * 2 * sn2 / alpha + s_dyn * mp.shape[1] * Q * (1 - alpha)
            gdyn, dx = self.dyn.aep_grads_mc(ms, vs, dmp, dvp, kfus, xs, alpha)
            gin_dyn = self.dyn.reparam(dx, vtm1, eps)
        else:
            logZ_dyn = s_dyn * np.sum(lz)
            dvt = s_dyn * (-0.5 / vsum + 0.5 * md**2 / vsum**2)
            dmt = s_dyn * (-md / vsum)
            dsn = np.sum(dvt) * 2 * sn2 / alpha + s_dyn * mp.shape[0] * Q * (1 - alpha)
            gdyn, gin_dyn = self.dyn.aep_grads_mm(mp, vp, -dmt, dvt, psi1, psi2, mtm1, vtm1, alpha)
        g = {'sn': dsn}
        # emission factors
        if self.gp_emi and mc:                               # aep_models.py:1127-1145
            sn_e = params['sn_emission']
            mo, vo, (ms, vs, kfus, xs, eps) = self.emi.prop_mc(mup, vup)
            lZe, dme, dve = gauss_log_Z_mc(sn_e, mo, vo, yb, alpha)
            logZ_emi = s_emi * lZe
            gemi, dx = self.emi.aep_grads_mc(ms, vs, s_emi * dme, s_emi * dve, kfus, xs, alpha)
            gin_emi = self.emi.reparam(dx, vup, eps)
            g['sn_emission'] = gauss_dsn_mc(sn_e, mo, dve, alpha, s_emi)
        elif self.gp_emi:
            sn_e = params['sn_emission']
            mo, vo, q1, q2 = self.emi.prop_mm(mup, vup)
            lZe, dme, dve = gauss_log_Z(sn_e, mo, vo, yb, alpha)
            logZ_emi = s_emi * lZe
            gemi, gin_emi = self.emi.aep_grads_mm(mo, vo, s_emi * dme, s_emi * dve,
                                                  q1, q2, mup, vup, alpha)
            g['sn_emission'] = gauss_dsn(sn_e, mo, dve, alpha, s_emi)
        else:
            logZ_emi, gin_emi, gemi = self.emi.tilted(mup, vup, alpha, s_emi, emi_idx)
        for k, val in gdyn.items():
            g[k + '_dynamic'] = val
        for k, val in gemi.items():
            g[k + '_emission'] = val
        dm_up, dv_up = gin_emi['mx'][:, :Q], gin_emi['vx'][:, :Q]
        dm_next, dv_next = gin_dyn['mx'][:, :Q], gin_dyn['vx'][:, :Q]
        # x gradients: three sources (aep_models.py:1163-1184, 1208-1315)
        one = np.ones((N, 1))
        s_post = -(1.0 - 1.0 / alpha) * one
        s_post[0:N - 1] += 1.0 / alpha
        s_post[1:N] += 1.0 / alpha
        p1, p2 = self.post1, self.post2
        gp1 = s_post * (p1 / p2)
        gp2 = s_post * (-0.5 * p1**2 / p2**2 - 0.5 / p2)
        gx1 = 3.0 * gp1
        gx2 = 6.0 * gp2 * self.f2
        gx1[[0, -1]] = 2.0 * gp1[[0, -1]]
        gx2[[0, -1]] = 4.0 * gp2[[0, -1]] * self.f2[[0, -1]]
        w = (3.0 - alpha) * one
        w[0] = 2.0 - alpha
        w[-1] = 2.0 - alpha
        sc = (-1.0 / alpha) * one
        sc[0:N - 1] += -1.0 / alpha
        sc[1:N] += -1.0 / alpha
        gx1 += sc * (cav1 / cav2) * w
        gx2 += sc * (-0.5 * cav1**2 / cav2**2 - 0.5 / cav2) * w * 2 * self.f2
        l1 = np.zeros_like(cav1)
        l2 = np.zeros_like(cav1)
        l1[emi_idx] = dm_up / cav2[emi_idx]
        l2[emi_idx] = -dm_up * cav1[emi_idx] / cav2[emi_idx]**2 - dv_up / cav2[emi_idx]**2
        ii = np.arange(emi_idx[0] + 1, emi_idx[-1] + 1)
        l1[ii] += dmt / cav2[ii]
        l2[ii] += -dmt * cav1[ii] / cav2[ii]**2 - dvt / cav2[ii]**2
        ii = np.arange(emi_idx[0], emi_idx[-1])
        l1[ii] += dm_next / cav2[ii]
        l2[ii] += -dm_next * cav1[ii] / cav2[ii]**2 - dv_next / cav2[ii]**2
        gx1 += l1 * w
        gx2 += l2 * w * 2 * self.f2
        g['x_factor_1'], g['x_factor_2'] = gx1, gx2
        # energy (aep_models.py:1186-1197, 1389-1437)
        m0, v0 = self.x_prior_1 / self.x_prior_2, 1.0 / self.x_prior_2
        phi_prior = 0.5 * Q * (m0**2 / v0 + np.log(v0))
        phi_post = np.sum(s_post * 0.5 * (p1**2 / p2 - np.log(p2)))
        phi_cav = np.sum(sc * 0.5 * (cav1**2 / cav2 - np.log(cav2)))
        energy = logZ_dyn + logZ_emi + phi_prior + phi_post + phi_cav + self.dyn.phi(alpha)
        if self.gp_emi:
            energy += self.emi.phi(alpha)
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


class VfeSGPSSM(_SSMBase):
    """vfe_models.py:866-1119."""

    def objective_function(self, params, mb_size, alpha='not_used', prop_mode=PROP_MM):
        N, Q = self.N, self.Din
        dyn_idx, emi_idx = self._window(mb_size)
        yb = self.y[emi_idx]
        nb = emi_idx.shape[0]
        s_dyn = -(N - 1) * 1.0 / dyn_idx.shape[0]
        s_emi = -N * 1.0 / nb
        self._set(params)
        pm, pv = self.post1[emi_idx] / self.post2[emi_idx], 1.0 / self.post2[emi_idx]
        mt, vt = pm[1:], pv[1:]
        mtm1, vtm1 = self._with_control(pm[:-1], pv[:-1], dyn_idx, self.Dcon_dyn)
        mup, vup = self._with_control(pm, pv, emi_idx, self.Dcon_emi)
        mc = prop_mode == PROP_MC
        sn2 = np.exp(2 * self.sn)                           # vfe_models.py:1080-1108
        if mc:                                              # vfe_models.py:983-995, 3-D branch 1094-1104
            mp, vp, (ms, vs, kfus, xs, eps) = self.dyn.prop_mc(mtm1, vtm1, cav=False)
            K = mp.shape[0]
            t2 = -0.5 / sn2 * (mt**2 + vt - 2 * mt * mp + mp**2 + vp)
            logZ_dyn = s_dyn * np.sum(-0.5 * np.log(2 * np.pi * sn2) + t2) / K
            dmp = s_dyn / sn2 * (mt - mp) / K
            dvp = -s_dyn * 0.5 / sn2 * np.ones_like(vp) / K
            dmt, dvt = -np.sum(dmp, axis=0), np.sum(dvp, axis=0)
            dsn = s_dyn * np.sum(-1 - 2 * t2) / K
            gdyn, dx = self.dyn.vfe_grads_mc(ms, vs, dmp, dvp, kfus, xs)
            gin_dyn = self.dyn.reparam(dx, vtm1, eps)
        else:
            mp, vp, psi1, psi2 = self.dyn.prop_mm(mtm1, vtm1, cav=False)
            t2 = -0.5 / sn2 * (mt**2 + vt - 2 * mt * mp + mp**2 + vp)
            logZ_dyn = s_dyn * np.sum(-0.5 * np.log(2 * np.pi * sn2) + t2)
            dmt = -s_dyn / sn2 * (mt - mp)
            dvt = -s_dyn * 0.5 / sn2 * np.ones_like(vt)
            dsn = s_dyn * np.sum(-1 - 2 * t2)
            gdyn, gin_dyn = self.dyn.vfe_grads_mm(mp, vp, -dmt, dvt, psi1, psi2, mtm1, vtm1)
        g = {'sn': dsn}
        if self.gp_emi and mc:                              # vfe_models.py:996-1010
            sn_e = params['sn_emission']
            mo, vo, (ms, vs, kfus, xs, eps) = self.emi.prop_mc(mup, vup, cav=False)
            K = mo.shape[0]
            sn2e = np.exp(2.0 * sn_e)
            lle = np.sum(np.mean(-0.5 * np.log(2 * np.pi * sn2e) - 0.5 * ((yb - mo)**2 + vo) / sn2e, axis=0))
            dme, dve = (yb - mo) / sn2e / K, -0.5 / sn2e * np.ones_like(vo) / K
            logZ_emi = s_emi * lle
            gemi, dx = self.emi.vfe_grads_mc(ms, vs, s_emi * dme, s_emi * dve, kfus, xs)
            gin_emi = self.emi.reparam(dx, vup, eps)
            g['sn_emission'] = s_emi * np.sum(-1 + ((yb - mo)**2 + vo) / sn2e) / K
        elif self.gp_emi:
            sn_e = params['sn_emission']
            mo, vo, q1, q2 = self.emi.prop_mm(mup, vup, cav=False)
            lle, dme, dve = gauss_log_lik_exp(sn_e, mo, vo, yb)
            logZ_emi = s_emi * lle
            gemi, gin_emi = self.emi.vfe_grads_mm(mo, vo, s_emi * dme, s_emi * dve,
                                                  q1, q2, mup, vup)
            g['sn_emission'] = gauss_dsn_log_lik_exp(sn_e, mo, vo, yb, s_emi)
        else:
            logZ_emi, gin_emi, gemi = self.emi.log_lik_exp(mup, vup, s_emi, emi_idx)
        for k, val in gdyn.items():
            g[k + '_dynamic'] = val
        for k, val in gemi.items():
            g[k + '_emission'] = val
        s_ent = -N * 1.0 / nb
        x_ent = s_ent * (nb * Q * (0.5 + 0.5 * np.log(2 * np.pi)) + np.sum(0.5 * np.log(pv)))
        dm = gin_emi['mx'][:, :Q].copy()
        dm[1:] += dmt
        dm[:-1] += gin_dyn['mx'][:, :Q]
        dv = gin_emi['vx'][:, :Q] + s_ent * 0.5 / pv
        dv[1:] += dvt
        dv[:-1] += gin_dyn['vx'][:, :Q]
        # base_models.py:1730-1752 compute_posterior_grad_x
        g1 = np.zeros_like(self.post1)
        g2 = np.zeros_like(self.post1)
        if self.nat_param:
            p1, p2 = self.post1[emi_idx], self.post2[emi_idx]
            a1 = dm / p2
            a2 = -dm * p1 / p2**2 - dv / p2**2
            sx = 3.0 * np.ones((nb, 1))
            sx[emi_idx == 0] = 2
            sx[emi_idx == N - 1] = 2
            g1[emi_idx] = a1 * sx
            g2[emi_idx] = a2 * sx * 2 * self.f2[emi_idx]
        else:
            g1[emi_idx] = dm
            g2[emi_idx] = dv * 2 * self.f2[emi_idx]
        g['x_factor_1'], g['x_factor_2'] = g1, g2
        energy = logZ_dyn + logZ_emi + x_ent + self.dyn.kl()
        if self.gp_emi:
            energy += self.emi.kl()
        for p in self.fixed_params:
            g[p] = np.zeros_like(g[p])
        return energy / N, {k: val / N for k, val in g.items()}


# --------------------------------------------------------------------------
# prediction (base_models.py:985-998, 1140-1158)
# --------------------------------------------------------------------------
def predict_sgpr(model, params, xs):
    L = model.layer
    L.set_params(params)
    m, v, _ = L.prop_det(xs, cav=False)
    return m, v


def predict_sdgpr(model, params, xs):
    for i, L in enumerate(model.layers):
        L.set_params(params, '_%d' % i)
    m, v, _ = model.layers[0].prop_det(xs, cav=False)
    for L in model.layers[1:]:
        m, v, _, _ = L.prop_mm(m, v, cav=False)
    return m, v
