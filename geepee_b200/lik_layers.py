"""Likelihood layers (reference: geepee/lik_layers.py).

Gauss_Layer (85-283): the per-row log-partition / expected log-likelihood and their
derivatives run in the `gauss_lik` CUDA kernel, fused with the scaling by scale_logZ.
Gauss_Emis (474-676): linear-Gaussian emission of the state-space model; a batched
Dout x Dout problem per row, evaluated on the device in fp64.
Probit_Layer (285-471): binary classification, closed form for
alpha = 1 and Gauss-Hermite quadrature otherwise, in the `probit_lik` CUDA kernel.
"""
import numpy as np
import torch

from . import ops
from . import tail as tl
from .layers import to_dev, default_device

_F = torch.float64


class Lik_Layer(object):
    has_sn = False          # does the layer own a noise hyper-parameter 'sn'?

    def __init__(self, N, D):
        self.N = N
        self.D = D

    def init_hypers(self, key_suffix=''):
        return {}

    def get_hypers(self, key_suffix=''):
        return {}

    def update_hypers(self, params, key_suffix=''):
        pass


class Probit_Layer(Lik_Layer):
    """lik_layers.py:285-471 (2-D branches; y in {-1,+1}; no hyper-parameters)."""

    def __init__(self, N, D, device=None):
        super(Probit_Layer, self).__init__(N, D)
        self.device = device = device if device is not None else default_device()
        from . import config
        gx, gw = np.polynomial.hermite.hermgauss(config.GH_DEGREE)   # lik_layers.py:290-301
        self._gx, self._gw = to_dev(gx, device), to_dev(gw, device)

    def update_hypers(self, params, key_suffix='', _dev=None):
        pass

    # ---- device path (same contract as Gauss_Layer; the 'sn' slot is a zero) ---------------
    def _log_Z(self, m, v, y, alpha, scale):
        dm, dv, o = ops.probit_lik(m, v, y, self._gx, self._gw, alpha, scale, 0)
        return dm, dv, o[0], o[1].reshape(())

    def _log_lik_exp(self, m, v, y, scale):
        dm, dv, o = ops.probit_lik(m, v, y, self._gx, self._gw, 1.0, scale, 1)
        return dm, dv, o[0], o[1].reshape(())

    def _log_Z_mc(self, m, v, y, alpha, scale):
        """lik_layers.py:364-409: m, v [K,n,D] from K Monte-Carlo samples; log-mean-exp over the
        samples of the per-sample tilted log-partitions (elementwise, on the device; the
        reference's eps = 1e-16 and its `+ eps` on the derivatives kept)."""
        eps = 1e-16
        yy = y.unsqueeze(0)
        if alpha == 1.0:
            t = yy * m / torch.sqrt(1 + v)
            Z = 0.5 * (1 + torch.erf(t / np.sqrt(2)))
            lt = torch.log(Z + eps)
        else:
            gx = self._gx.reshape(-1, 1, 1, 1)
            gw = self._gw.reshape(-1, 1, 1, 1)
            ts = gx * torch.sqrt(2 * v) + m
            pdfs = 0.5 * (1 + torch.erf(yy * ts / np.sqrt(2))) + eps
            Zt = (pdfs**alpha * gw).sum(0) / np.sqrt(np.pi)
            lt = torch.log(Zt)
        lmax = lt.max(dim=0).values
        ex = torch.exp(lt - lmax)
        se = ex.sum(0)
        logZ = (lmax + torch.log(se) - np.log(m.shape[0])).sum()
        w = ex / se
        if alpha == 1.0:
            dt = 1 / (Z + eps) / np.sqrt(2 * np.pi) * torch.exp(-t**2 / 2)
            dm = w * dt * yy / torch.sqrt(1 + v)
            dv = w * dt * (-0.5 * yy * m / (1 + v)**1.5)
        else:
            a = pdfs**(alpha - 1.0) * torch.exp(-ts**2 / 2)
            dm = w * ((gw * a).sum(0) * yy * alpha / np.pi / np.sqrt(2)) / Zt + eps
            dv = w * ((gw * (a * gx)).sum(0) * yy * alpha / np.pi / np.sqrt(2) / torch.sqrt(2 * v)) / Zt + eps
        zero = torch.zeros((), dtype=_F, device=m.device)
        return (scale * dm).contiguous(), (scale * dv).contiguous(), logZ, zero

    # ---- reference API (numpy) ---------------------------------------------------------------
    def compute_log_Z(self, mout, vout, y, alpha=1.0, compute_dm2=False):
        """lik_layers.py:303-411 (2-D and Monte-Carlo 3-D branches)."""
        if mout.ndim not in (2, 3):
            raise RuntimeError('invalid ndim, ndim=%d' % mout.ndim)
        if compute_dm2:
            raise NotImplementedError('dm2 of the probit likelihood is used by pep_models only (out of scope)')
        dev = self.device
        if mout.ndim == 3:
            dm, dv, logZ, _ = self._log_Z_mc(to_dev(mout, dev), to_dev(vout, dev), to_dev(y, dev), alpha, 1.0)
            return float(logZ.item()), dm.cpu().numpy(), dv.cpu().numpy()
        dm, dv, o = ops.probit_lik(to_dev(mout, dev), to_dev(vout, dev), to_dev(y, dev), self._gx, self._gw,
                                   alpha, 1.0, 0)
        return float(o[0].item()), dm.cpu().numpy(), dv.cpu().numpy()

    def compute_log_lik_exp(self, m, v, y):
        """lik_layers.py:418-457: 3-D inputs (Monte-Carlo propagation) are averaged over the K samples."""
        dev = self.device
        if m.ndim == 3:
            K = m.shape[0]
            dm, dv, o = ops.probit_lik(to_dev(m, dev).reshape(-1, self.D), to_dev(v, dev).reshape(-1, self.D),
                                       to_dev(y, dev).repeat(K, 1), self._gx, self._gw, 1.0, 1.0 / K, 1)
            return float(o[0].item()) / K, dm.reshape(m.shape).cpu().numpy(), dv.reshape(m.shape).cpu().numpy()
        dm, dv, o = ops.probit_lik(to_dev(m, dev), to_dev(v, dev), to_dev(y, dev), self._gx, self._gw,
                                   1.0, 1.0, 1)
        return float(o[0].item()), dm.cpu().numpy(), dv.cpu().numpy()

    def backprop_grads(self, mout, vout, dmout, dvout, alpha=1.0, scale=1.0):
        return {}

    def backprop_grads_log_lik_exp(self, m, v, dm, dv, y, scale=1.0):
        return {}

    def output_probabilistic(self, mf, vf, alpha=1.0):
        raise NotImplementedError('TODO: return probablity of y=1')     # lik_layers.py:471


class Gauss_Layer(Lik_Layer):
    has_sn = True

    def __init__(self, N, D, device=None):
        super(Gauss_Layer, self).__init__(N, D)
        self.sn = 0
        self.device = device if device is not None else default_device()
        self._sn = None

    # ---- device path ------------------------------------------------------------------------
    def _log_Z(self, m, v, y, alpha, scale):
        """lik_layers.py:104-133 + 154-181.  Returns (scale*dm, scale*dv, logZ_sum, dsn) with
        logZ_sum unscaled and dsn = scale*(sum(dv) 2 sn2/alpha + n D (1-alpha)), all on device."""
        dm, dv, o = ops.gauss_lik(m, v, y, self._sn, alpha, scale, 0)
        sn2 = float(np.exp(2.0 * np.ravel(self.sn)[0]))           # host copy of the same parameter
        dsn = tl.dots([(scale * 2.0 * sn2 / alpha, o[1:2], None)], const=scale * m.shape[0] * self.D * (1.0 - alpha))
        return dm, dv, o[0], dsn.reshape(())

    def _log_Z_mc(self, m, v, y, alpha, scale):
        """lik_layers.py:134-150 + 175-181: m, v [K,n,D] from K Monte-Carlo samples of the layer
        input; log-mean-exp of the per-sample tilted log-partitions (elementwise, on the device).
        Same contract as _log_Z: (scale*dm, scale*dv, logZ_sum unscaled, dsn)."""
        sn2 = torch.exp(2.0 * self._sn)
        vv = v + sn2 / alpha
        d = y.unsqueeze(0) - m
        lz = -0.5 * (torch.log(2 * np.pi * vv) + d * d / vv) \
            + (0.5 * torch.log(2 * np.pi * sn2 / alpha) - 0.5 * alpha * torch.log(2 * np.pi * sn2))
        lmax = lz.max(dim=0).values
        ex = torch.exp(lz - lmax)
        se = ex.sum(0)
        logZ = (lmax + torch.log(se) - np.log(m.shape[0])).sum()
        w = ex / se
        dm = w * d / vv
        dv = w * (-0.5 / vv + 0.5 * d * d / (vv * vv))
        dsn = scale * (dv.sum() * 2.0 * sn2 / alpha + m.shape[1] * self.D * (1.0 - alpha))
        return (scale * dm).contiguous(), (scale * dv).contiguous(), logZ, dsn.reshape(())

    def _log_lik_exp(self, m, v, y, scale):
        """lik_layers.py:183-199 + 217-226."""
        dm, dv, o = ops.gauss_lik(m, v, y, self._sn, 1.0, scale, 1)
        return dm, dv, o[0], tl.dots([(scale, o[1:2], None)]).reshape(())

    # ---- reference API (numpy) ---------------------------------------------------------------
    def compute_log_Z(self, mout, vout, y, alpha=1.0, compute_dm2=False):
        """lik_layers.py:104-152: 2-D branch and the 3-D branch of Monte-Carlo propagation (log-mean-exp
        over the K samples).  Like the reference, adds sn2/alpha to the caller's vout in place (122/136)."""
        if mout.ndim not in (2, 3):
            raise RuntimeError('invalid ndim, ndim=%d' % mout.ndim)
        dev = self._sn.device
        if mout.ndim == 3:
            dm, dv, logZ, _ = self._log_Z_mc(to_dev(mout, dev), to_dev(vout, dev), to_dev(y, dev), alpha, 1.0)
            vout += np.exp(2.0 * self.sn) / alpha
            return float(logZ.item()), dm.cpu().numpy(), dv.cpu().numpy()
        dm, dv, o = ops.gauss_lik(to_dev(mout, dev), to_dev(vout, dev), to_dev(y, dev), self._sn, alpha, 1.0, 0)
        vout += np.exp(2.0 * self.sn) / alpha
        if compute_dm2:
            return float(o[0].item()), dm.cpu().numpy(), dv.cpu().numpy(), -1.0 / vout
        return float(o[0].item()), dm.cpu().numpy(), dv.cpu().numpy()

    def backprop_grads(self, mout, vout, dmout, dvout, alpha=1.0, scale=1.0):
        """lik_layers.py:154-181."""
        if mout.ndim not in (2, 3):
            raise RuntimeError('invalid ndim, ndim=%d' % mout.ndim)
        sn2 = np.exp(2.0 * self.sn)
        dim_prod = mout.shape[mout.ndim - 2] * self.D
        return {'sn': scale * (np.sum(dvout) * 2 * sn2 / alpha + dim_prod * (1 - alpha))}

    def compute_log_lik_exp(self, mout, vout, y):
        """lik_layers.py:183-218: expected log-likelihood; 3-D inputs are averaged over the K samples."""
        if mout.ndim not in (2, 3):
            raise RuntimeError('invalid ndim, ndim=%d' % mout.ndim)
        dev = self._sn.device
        if mout.ndim == 3:
            K = mout.shape[0]
            yk = to_dev(y, dev).repeat(K, 1)
            dm, dv, o = ops.gauss_lik(to_dev(mout, dev).reshape(-1, self.D), to_dev(vout, dev).reshape(-1, self.D),
                                      yk, self._sn, 1.0, 1.0 / K, 1)
            return float(o[0].item()) / K, dm.reshape(mout.shape).cpu().numpy(), dv.reshape(mout.shape).cpu().numpy()
        dm, dv, o = ops.gauss_lik(to_dev(mout, dev), to_dev(vout, dev), to_dev(y, dev), self._sn, 1.0, 1.0, 1)
        return float(o[0].item()), dm.cpu().numpy(), dv.cpu().numpy()

    def backprop_grads_log_lik_exp(self, m, v, dm, dv, y, scale=1.0):
        """lik_layers.py:220-236: d/dsn of the expected log-likelihood (the gauss_lik kernel's second
        output is sum(-1 + ((y-m)^2+v)/sn2))."""
        if m.ndim not in (2, 3):
            raise RuntimeError('invalid ndim, ndim=%d' % m.ndim)
        dev = self._sn.device
        K = m.shape[0] if m.ndim == 3 else 1
        yk = to_dev(y, dev).repeat(K, 1) if m.ndim == 3 else to_dev(y, dev)
        _, _, o = ops.gauss_lik(to_dev(m, dev).reshape(-1, self.D), to_dev(v, dev).reshape(-1, self.D), yk,
                                self._sn, 1.0, 1.0, 1)
        return {'sn': scale * float(o[1].item()) / K}

    def output_probabilistic(self, mf, vf, alpha=1.0):
        """lik_layers.py:238-249."""
        return mf, vf + np.exp(2.0 * self.sn) / alpha

    def init_hypers(self, key_suffix=''):
        self.sn = np.log(0.01)
        return {'sn' + key_suffix: self.sn}

    def get_hypers(self, key_suffix=''):
        return {'sn' + key_suffix: self.sn}

    def update_hypers(self, params, key_suffix='', _dev=None):
        self.sn = params['sn' + key_suffix]
        if _dev is not None:
            self._sn = _dev['sn' + key_suffix].reshape(-1)[:1].contiguous()
        else:
            self._sn = to_dev(np.reshape(self.sn, (-1,))[:1], self.device)


class Gauss_Emis(object):
    """lik_layers.py:474-676: y ~ N(C x, diag(R))."""

    def __init__(self, y, Dout, Din, device=None):
        self.y = y
        self.N = y.shape[0]
        self.Dout = Dout
        self.Din = Din
        self.device = device = device if device is not None else default_device()
        self.C = np.zeros((Dout, Din))
        self.R = np.zeros(Dout)
        self._y = to_dev(y, device)

    def update_hypers(self, params, key_suffix='', _dev=None):
        self.C = params['C' + key_suffix]
        self.R = np.exp(2 * params['R' + key_suffix])
        if _dev is not None:
            self._C = _dev['C' + key_suffix].reshape(self.Dout, self.Din).contiguous()
            self._R = torch.exp(2.0 * _dev['R' + key_suffix].reshape(self.Dout))
        else:
            self._C = to_dev(self.C, self.device)
            self._R = to_dev(self.R, self.device)

    def init_hypers(self, key_suffix=''):
        return {'C' + key_suffix: np.ones((self.Dout, self.Din)) / (self.Dout * self.Din),
                'R' + key_suffix: np.log(0.01) * np.ones(self.Dout)}

    def get_hypers(self, key_suffix=''):
        return {'C' + key_suffix: self.C, 'R' + key_suffix: 0.5 * np.log(self.R)}

    def output_probabilistic(self, mf, vf):
        my = np.einsum('ab,nb->na', self.C, mf)
        vy_noiseless = np.einsum('ab,nb,bc->nac', self.C, vf, self.C.T)
        return my, vy_noiseless, vy_noiseless + np.diag(self.R)

    def _tilted(self, mx, vx, alpha, scale, y):
        """lik_layers.py:573-627 on the device: per-row Dout x Dout Cholesky.
        Returns (scale*logZ, scale*dmx, scale*dvx, {'C','R'} grads)."""
        C, R, Do = self._C, self._R, self.Dout
        Nb = mx.shape[0]
        if ops.gauss_emis_supported(Do, self.Din) and Nb > 0:
            # fused per-row kernel (Cholesky / inverse of the Do x Do matrices in registers)
            dmx, dvx, out = ops.gauss_emis(mx.contiguous(), vx.contiguous(), y.contiguous(), C, R, alpha, scale)
            fin = ops.gauss_emis_finish(out, R, alpha, scale, Nb, Do, self.Din)     # lik_layers.py:600-627
            return fin[0], dmx, dvx, {'C': fin[2 + Do:].reshape(Do, self.Din), 'R': fin[2:2 + Do]}
        CVC = torch.einsum('da,na,ba->ndb', C, vx, C)
        Vy = torch.diag(R / alpha).unsqueeze(0) + CVC
        Yd = y - torch.matmul(mx, C.t())
        Lc, _ = torch.linalg.cholesky_ex(Vy, check_errors=False)   # no host sync
        VinvY = torch.cholesky_solve(Yd.unsqueeze(-1), Lc).squeeze(-1)
        quad = -0.5 * (Yd * VinvY).sum()
        # log|I + alpha CVC / R| = log|Vy| - sum log(R/alpha)
        ld_Vy = 2.0 * torch.log(torch.diagonal(Lc, dim1=1, dim2=2)).sum()
        vlog = -0.5 * (ld_Vy - Nb * torch.log(R / alpha).sum())
        logZ = (-Nb * Do * 0.5 * alpha * np.log(2 * np.pi) - 0.5 * Nb * alpha * torch.log(R).sum()
                + vlog + quad)
        Vyinv = torch.cholesky_inverse(Lc)
        dR = (-0.5 * torch.diagonal(Vyinv, dim1=1, dim2=2).sum(0) + 0.5 * (VinvY**2).sum(0)) / alpha
        dR = (dR + 0.5 * Nb * (1 - alpha) / R) * 2 * R
        dSig = -0.5 * Vyinv + 0.5 * VinvY.unsqueeze(-1) * VinvY.unsqueeze(-2)
        dC = torch.matmul(VinvY.t(), mx) + 2.0 * torch.einsum('nc,bc,nab->ac', vx, C, dSig)
        dmx = torch.matmul(VinvY, C)
        dvx = torch.einsum('nab,ad,bd->nd', dSig, C, C)
        return logZ * scale, dmx * scale, dvx * scale, {'C': dC * scale, 'R': dR * scale}

    def _log_lik_exp(self, mx, vx, scale, y):
        """lik_layers.py:629-676 on the device."""
        C, R, Do = self._C, self._R, self.Dout
        Nb = mx.shape[0]
        Cm = torch.matmul(mx, C.t())
        CRC_diag = (C * C / R.unsqueeze(1)).sum(0)
        sv = vx.sum(0)
        res2 = ((y - Cm)**2).sum(0)
        logZ = (-0.5 * Nb * Do * np.log(2 * np.pi) - 0.5 * Nb * torch.log(R).sum()
                - 0.5 * (res2 / R).sum() - 0.5 * (sv * CRC_diag).sum())
        dR = (-0.5 * Nb / R + 0.5 * res2 / R**2 + 0.5 * (C * C * sv.unsqueeze(0)).sum(1) / R**2) * 2 * R
        dC = torch.matmul((y - Cm).t(), mx) / R.unsqueeze(1) - C * sv.unsqueeze(0) / R.unsqueeze(1)
        dmx = torch.matmul((y - Cm) / R.unsqueeze(0), C)
        dvx = (-0.5 * CRC_diag).unsqueeze(0).expand(Nb, -1).contiguous()
        return logZ * scale, dmx * scale, dvx * scale, {'C': dC * scale, 'R': dR * scale}

    # reference API (numpy)
    def compute_emission_tilted(self, mx, vx, alpha, scale, idxs=None):
        if idxs is None:
            idxs = np.arange(self.N)
        dev = self.device
        y = self._y[torch.as_tensor(idxs, device=dev)]
        lz, dmx, dvx, g = self._tilted(to_dev(mx, dev), to_dev(vx, dev), alpha, scale, y)
        return (float(lz.item()), {'mx': dmx.cpu().numpy(), 'vx': dvx.cpu().numpy()},
                {k: v.cpu().numpy() for k, v in g.items()})

    def compute_emission_log_lik_exp(self, mx, vx, scale, idxs=None):
        if idxs is None:
            idxs = np.arange(self.N)
        dev = self.device
        y = self._y[torch.as_tensor(idxs, device=dev)]
        lz, dmx, dvx, g = self._log_lik_exp(to_dev(mx, dev), to_dev(vx, dev), scale, y)
        return (float(lz.item()), {'mx': dmx.cpu().numpy(), 'vx': dvx.cpu().numpy()},
                {k: v.cpu().numpy() for k, v in g.items()})
