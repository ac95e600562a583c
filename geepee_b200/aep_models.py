"""AEP (black-box alpha / approximate EP) models on the B200 hot path.

Reference: geepee/aep_models.py -- SGPR 589-667, SGPLVM 670-867, SDGPR 870-988,
SGPSSM 991-1437, SDGPR_H 1440-1864.  Same constructors, same ``objective_function(params, mb_size, alpha,
prop_mode) -> (energy, grads)`` with the same dict keys and shapes.

Every objective has the same three-phase shape:
  1. per-row phase on this rank's slice of the minibatch: fused CUDA kernels ->
     additive sufficient statistics (+ per-row input gradients that stay local);
  2. one packed all-reduce of the statistics (no-op on one GPU);
  3. the replicated O(Dout M^3) tail in fp64 -> gradients wrt every parameter.
"""
import numpy as np
import torch

from . import dist, nvtx, ops
from . import tail as tl
from .sched import TailStreams
from .base_models import Base_SGPR, Base_SDGPR, Base_SGPLVM, Base_SGPSSM
from .config import PROP_MM, PROP_MC, PROP_LIN, MC_NO_SAMPLES
from .layers import pack_to_device, to_dev, AEP_SGP_Layer as SGP_Layer  # noqa: F401  (reference name)

_F = torch.float64


def _check_mode(prop_mode, mc_ok=False):
    if prop_mode == PROP_MM or (mc_ok and prop_mode == PROP_MC):
        return
    if prop_mode in (PROP_MC, PROP_LIN):
        raise NotImplementedError('prop_mode %s: not part of the B200 hot path yet '
                                  '(SURVEY.md section 8f)' % prop_mode)
    raise NotImplementedError('propagation mode not implemented')


def _mc_eps(n, Q, dev):
    """eps[K, n, Q] of the Monte-Carlo propagation: drawn on the host from numpy's GLOBAL RNG
    exactly where the reference draws it (aep_models.py:171, base_models.py:320), so that seeded
    runs reproduce the reference; every rank draws the same array and keeps its rows."""
    eps = dist.agree(np.random.randn(MC_NO_SAMPLES, n, Q))
    lo, hi = dist.shard(n)
    return to_dev(eps[:, lo:hi], dev)


def _zeros(dev, *shape):
    return torch.zeros(shape, dtype=_F, device=dev)


def _zero_stats(layer, with_rows=0):
    dev, Do, M, Q = layer.device, layer.Dout, layer.M, layer.Din
    st = {'dA': _zeros(dev, Do, M), 'dB': _zeros(dev, Do, M, M), 'dzu': _zeros(dev, M, Q),
          'dl': _zeros(dev, Q), 'dsf2': _zeros(dev, 1), 'dvsum': _zeros(dev, 1)}
    return st


_STAT_KEYS = ('dA', 'dB', 'dzu', 'dl', 'dsf2', 'dvsum')


def _add_stats(add, prefix, st):
    for k in _STAT_KEYS:
        add[prefix + k] = st[k]


def _get_stats(add, prefix):
    return {k: add[prefix + k] for k in _STAT_KEYS}


class SGPR(Base_SGPR):
    """aep_models.py:589-667."""

    def __init__(self, x_train, y_train, no_pseudo, lik='Gaussian', nat_param=True,
                 prec=None, device=None):
        super(SGPR, self).__init__(x_train, y_train, no_pseudo, lik, nat_param, prec, device)
        self.sgp_layer = SGP_Layer(self.N, self.Din, self.Dout, self.M, nat_param, prec, self.device)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        N, L, dev = self.N, self.sgp_layer, self.device
        xb, yb, n = self._batch(mb_size)
        scale_logZ = -N * 1.0 / n / alpha
        L._fuse_cavity_alpha = alpha
        self.update_hypers(params)
        L.compute_cavity(alpha)
        add = {}
        if xb.shape[0] > 0:
            def lik(m, v, c0, c1):
                dm, dv, logZ, dsn = self.lik_layer._log_Z(m, v, yb[c0:c1], alpha, scale_logZ)
                return dm, dv, {'logZ': logZ.reshape(1), 'dsn': dsn.reshape(1)}
            st, ext = L.det_step(xb, lik, cav=True)       # row-chunked: saved Kfu / T stay bounded
            _add_stats(add, 's_', st)
            add.update(ext)
        else:
            _add_stats(add, 's_', _zero_stats(L))
            add['logZ'], add['dsn'] = _zeros(dev, 1), _zeros(dev, 1)
        add = dist.allreduce_dict(add)
        grads = L._tail_det(_get_stats(add, 's_'), alpha)
        if self.lik_layer.has_sn:
            grads['sn'] = add['dsn'].reshape(())
        energy = tl.dots([(scale_logZ, add['logZ'], None), (1.0, L._phi(alpha), None)])
        return self._finish(energy, grads)


class SDGPR(Base_SDGPR):
    """aep_models.py:870-988 (layers are always natural-parameter: line 893)."""

    def __init__(self, x_train, y_train, no_pseudos, hidden_sizes, lik='Gaussian',
                 prec=None, device=None):
        super(SDGPR, self).__init__(x_train, y_train, no_pseudos, hidden_sizes, lik, prec, device)
        self.sgp_layers = [SGP_Layer(self.N, self.size[i], self.size[i + 1], self.Ms[i], True,
                                     prec, self.device) for i in range(self.L)]
        self._tail_streams = TailStreams(self.device, self.L)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        """aep_models.py:895-988.  Per-row kernels on the current stream; every layer's tail (and,
        with several ranks, the all-reduce of its statistics) on its own side stream (sched.py)."""
        _check_mode(prop_mode)
        N, dev = self.N, self.device
        xb, yb, n = self._batch(mb_size)
        scale_logZ = -N * 1.0 / n / alpha
        ts = self._tail_streams
        pdev = pack_to_device(params, dev)
        self.lik_layer.update_hypers(params, _dev=pdev)
        uploaded = ts.mark()

        def pre_tail(i):
            # q(u) + cavity of layer i on its side stream; issued right after the forward kernel
            # of layer i-1 was queued, so it runs underneath that kernel
            layer = self.sgp_layers[i]
            layer._fuse_cavity_alpha = alpha
            ts.fork(i, after=uploaded)
            with ts.on(i):
                layer.update_hypers(params, key_suffix='_%d' % i, _dev=pdev)
                layer.compute_cavity(alpha)

        grads, phis = {}, [None] * self.L
        has_rows = xb.shape[0] > 0
        ctxs = []
        pre_tail(0)
        if has_rows:
            m = v = None
            for i, layer in enumerate(self.sgp_layers):
                ts.join(i)
                if i == 0:
                    m, v, ctx = layer._fwd_det(xb, cav=True, save=True)
                else:
                    m, v, ctx = layer._fwd_mm(m, v, cav=True)
                ctxs.append(ctx)
                if i + 1 < self.L:
                    pre_tail(i + 1)
            dmi, dvi, logZ, dsn = self.lik_layer._log_Z(m, v, yb, alpha, scale_logZ)
        else:
            for i in range(1, self.L):
                pre_tail(i)
            ts.join_all()
        top = None
        for i in range(self.L - 1, -1, -1):
            layer = self.sgp_layers[i]
            add = {}
            if has_rows:
                if i == 0:
                    st = layer._bwd_det(ctxs[0], dmi, dvi)
                else:
                    st = layer._bwd_mm(ctxs[i], dmi, dvi)
                    dmi, dvi = st['dmx'], st['dvx']
                _add_stats(add, 's_', st)
            else:
                _add_stats(add, 's_', _zero_stats(layer))
            if i == self.L - 1:
                add['logZ'] = logZ.reshape(1) if has_rows else _zeros(dev, 1)
                add['dsn'] = dsn.reshape(1) if has_rows else _zeros(dev, 1)
            ts.fork(i)
            ts.keep(i, add.values())
            with ts.on(i):
                add = dist.allreduce_dict(add)
                g = layer._tail_det(_get_stats(add, 's_'), alpha) if i == 0 else \
                    layer._tail_mm(_get_stats(add, 's_'), alpha)
                for k, val in g.items():
                    grads[k + '_%d' % i] = val
                phis[i] = layer._phi(alpha)
                if i == self.L - 1:
                    top = add
        ts.join_all()
        energy = tl.dots([(scale_logZ, top['logZ'], None)] + [(1.0, phis[i], None) for i in range(self.L)])
        if self.lik_layer.has_sn:
            grads['sn'] = top['dsn'].reshape(())
        return self._finish(energy, grads)


class SDGPR_H(Base_SDGPR):
    """aep_models.py:1440-1864: deep GP regression with inference for the hidden variables -- one
    Gaussian factor per training row and hidden unit (`h_factor_1/2_<i>[N, size[i+1]]`, tied
    twice: posterior = 2 x factor, cavity = (2 - alpha) x factor) and a transition noise
    `sn_hidden[i]` per hidden layer.

    The hidden variables decouple the layers: layer i propagates the cavity of hidden layer i-1
    (the inputs for i = 0) and is matched against the cavity of hidden layer i
    (compute_transition_tilted, 1710-1745), so every layer's forward + backward pair runs back to
    back on this rank's rows and only the additive statistics meet in the all-reduce.  Full batch
    only, like the reference (compute_grads_hidden, 1621-1653, combines the [N, D] factors with
    batch-sized gradients)."""

    def __init__(self, x_train, y_train, no_pseudos, hidden_sizes, lik='Gaussian',
                 prec=None, device=None):
        super(SDGPR_H, self).__init__(x_train, y_train, no_pseudos, hidden_sizes, lik, prec, device)
        self.sgp_layers = [SGP_Layer(self.N, self.size[i], self.size[i + 1], self.Ms[i], True,
                                     prec, self.device) for i in range(self.L)]
        self.sn = np.zeros(self.L - 1)
        self.h_factor_1 = [np.zeros((self.N, self.size[i + 1])) for i in range(self.L - 1)]
        self.h_factor_2 = [np.zeros((self.N, self.size[i + 1])) for i in range(self.L - 1)]

    # ---- parameters (aep_models.py:1808-1864) -------------------------------------------------
    def init_hypers(self, y_train):
        init_params = super(SDGPR_H, self).init_hypers(y_train)
        init_params['sn_hidden'] = np.log(0.001) * np.ones(self.L - 1)
        for i in range(self.L - 1):
            init_params['h_factor_1_%d' % i] = np.zeros((self.N, self.size[i + 1]))
            init_params['h_factor_2_%d' % i] = np.log(0.01) * np.ones((self.N, self.size[i + 1]))
        return init_params

    def get_hypers(self):
        params = super(SDGPR_H, self).get_hypers()
        params['sn_hidden'] = self.sn
        for i in range(self.L - 1):
            params['h_factor_1_%d' % i] = self.h_factor_1[i]
            params['h_factor_2_%d' % i] = np.log(self.h_factor_2[i]) / 2
        return params

    def update_hypers(self, params, _dev=None):
        dev = pack_to_device(params, self.device) if _dev is None else _dev
        for i, layer in enumerate(self.sgp_layers):
            layer.update_hypers(params, key_suffix='_%d' % i, _dev=dev)
        self.lik_layer.update_hypers(params, _dev=dev)
        self.sn = params['sn_hidden']
        for i in range(self.L - 1):
            self.h_factor_1[i] = params['h_factor_1_%d' % i]
            self.h_factor_2[i] = np.exp(2 * np.asarray(params['h_factor_2_%d' % i]))

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        _check_mode(prop_mode)
        N, dev, Ln = self.N, self.device, self.L
        if mb_size < N:
            raise NotImplementedError('SDGPR_H: full batches only (the reference combines the [N, D] '
                                      'hidden factors with batch-sized gradients, aep_models.py:1621-1653)')
        lo, hi = dist.shard(N)
        xb, yb = self._x[lo:hi], self._y[lo:hi]
        scale = -1.0 / alpha                     # scale_logZ = -N / batch_size / alpha, full batch
        s_post, s_cav = -(1.0 - 1.0 / alpha), -1.0 / alpha
        for layer in self.sgp_layers:
            layer._fuse_cavity_alpha = alpha
        pdev = pack_to_device(params, dev)
        self.update_hypers(params, _dev=pdev)
        for layer in self.sgp_layers:
            layer.compute_cavity(alpha)
        sn2 = torch.exp(2.0 * pdev['sn_hidden'].reshape(-1))
        # cavity of the hidden variables of this rank's rows (compute_cavity_h, 1747-1767)
        h1 = [pdev['h_factor_1_%d' % i].reshape(N, -1)[lo:hi] for i in range(Ln - 1)]
        h2 = [torch.exp(2.0 * pdev['h_factor_2_%d' % i].reshape(N, -1)[lo:hi]) for i in range(Ln - 1)]
        c1 = [a * (2.0 - alpha) for a in h1]
        c2 = [a * (2.0 - alpha) for a in h2]
        cm = [(a / b).contiguous() for a, b in zip(c1, c2)]
        cv = [(1.0 / b).contiguous() for b in c2]
        add = {'logZ': _zeros(dev, 1), 'dsn_hidden': _zeros(dev, Ln - 1), 'dsn': _zeros(dev, 1),
               'phi_h': _zeros(dev, 1)}
        dmc, dvc = [], []
        has_rows = hi > lo
        for i, layer in enumerate(self.sgp_layers):
            if not has_rows:
                _add_stats(add, 's%d_' % i, _zero_stats(layer))
                continue
            if i == 0:
                mp, vp, ctx = layer._fwd_det(xb, cav=True, save=True)
            else:
                mp, vp, ctx = layer._fwd_mm(cm[i - 1], cv[i - 1], cav=True)
            if i < Ln - 1:
                # compute_transition_tilted (1710-1745) against the cavity of hidden layer i
                vsum = cv[i] + vp + sn2[i] / alpha
                md = cm[i] - mp
                lz = (-0.5 * md**2 / vsum - 0.5 * torch.log(2 * np.pi * vsum)).sum() \
                    + (hi - lo) * self.size[i + 1] * (0.5 * (1 - alpha) * torch.log(2 * np.pi * sn2[i])
                                                      - 0.5 * np.log(alpha))
                dvt = scale * (-0.5 / vsum + 0.5 * md**2 / vsum**2)
                dmt = scale * (-md / vsum)
                add['logZ'] = add['logZ'] + scale * lz
                add['dsn_hidden'][i] = dvt.sum() * 2 * sn2[i] / alpha \
                    + scale * (hi - lo) * self.size[i + 1] * (1 - alpha)
                dmp, dvp = (-dmt).contiguous(), dvt.contiguous()
            else:
                dmp, dvp, lzl, dsn = self.lik_layer._log_Z(mp, vp, yb, alpha, scale)
                add['logZ'] = add['logZ'] + scale * lzl
                add['dsn'] = dsn.reshape(1)
            if i == 0:
                st = layer._bwd_det(ctx, dmp, dvp)
            else:
                st = layer._bwd_mm(ctx, dmp, dvp)
                dmc[i - 1] = dmc[i - 1] + st['dmx']
                dvc[i - 1] = dvc[i - 1] + st['dvx']
            _add_stats(add, 's%d_' % i, st)
            if i < Ln - 1:
                dmc.append(dmt)
                dvc.append(dvt)
        # compute_grads_hidden (1621-1653) + compute_phi_{cavity,posterior}_h (1655-1690), this
        # rank's rows; the [N, D] gradients are assembled by the same all-reduce
        for i in range(Ln - 1):
            g1 = _zeros(dev, N, self.size[i + 1])
            g2 = _zeros(dev, N, self.size[i + 1])
            if has_rows:
                p1, p2 = 2.0 * h1[i], 2.0 * h2[i]
                d1 = (2 - alpha) * (dmc[i] / c2[i]) + s_cav * (2 - alpha) * (c1[i] / c2[i]) \
                    + s_post * 2 * (p1 / p2)
                d2 = (2 - alpha) * (-dmc[i] * c1[i] / c2[i]**2 - dvc[i] / c2[i]**2) \
                    + s_cav * (2 - alpha) * (-0.5 * c1[i]**2 / c2[i]**2 - 0.5 / c2[i]) \
                    + s_post * (-p1**2 / p2**2 - 1 / p2)
                g1[lo:hi] = d1
                g2[lo:hi] = 2 * d2 * h2[i]
                add['phi_h'] = add['phi_h'] + s_cav * (0.5 * (c1[i]**2 / c2[i] - torch.log(c2[i]))).sum() \
                    + s_post * (0.5 * (p1**2 / p2 - torch.log(p2))).sum()
            add['gh1_%d' % i], add['gh2_%d' % i] = g1, g2
        add = dist.allreduce_dict(add)
        grads = {}
        energy = add['logZ'] + add['phi_h']
        for i, layer in enumerate(self.sgp_layers):
            st = _get_stats(add, 's%d_' % i)
            g = layer._tail_det(st, alpha) if i == 0 else layer._tail_mm(st, alpha)
            for k, val in g.items():
                grads[k + '_%d' % i] = val
            energy = energy + layer._phi(alpha)
            if i < Ln - 1:
                grads['h_factor_1_%d' % i], grads['h_factor_2_%d' % i] = add['gh1_%d' % i], add['gh2_%d' % i]
        if self.lik_layer.has_sn:
            grads['sn'] = add['dsn'].reshape(())
        grads['sn_hidden'] = add['dsn_hidden']
        return self._finish(energy, grads)


class SGPLVM(Base_SGPLVM):
    """aep_models.py:670-867."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 nat_param=True, prec=None, device=None):
        super(SGPLVM, self).__init__(y_train, hidden_size, no_pseudo, lik, prior_mean, prior_var,
                                     nat_param, prec, device)
        self.sgp_layer = SGP_Layer(self.N, self.Din, self.Dout, self.M, nat_param, prec, self.device)

    def get_cavity_x(self, alpha, idxs=None):
        """aep_models.py:840-861 (numpy API)."""
        if idxs is None:
            idxs = np.arange(self.N)
        sel = torch.as_tensor(idxs, device=self.device)
        m, v = self._cavity_x(alpha, sel)
        return m.cpu().numpy(), v.cpu().numpy()

    def _cavity_x(self, alpha, sel):
        """numpy-API twin of ops.lvm_x_fwd (the objective uses the kernel)."""
        if self.nat_param:
            c1 = self.prior_x1 + (1.0 - alpha) * self._f1[sel]
            c2 = self.prior_x2 + (1.0 - alpha) * self._f2[sel]
        else:
            mpost, vpost = self._f1[sel], self._f2[sel]
            c1 = self.prior_x1 + (mpost / vpost - self.prior_x1) * (1 - alpha)
            c2 = self.prior_x2 + (1 / vpost - self.prior_x2) * (1 - alpha)
        return (c1 / c2).contiguous(), (1.0 / c2).contiguous()

    @staticmethod
    def _phi_x(mx, vx):
        """aep_models.py:863-867."""
        return (0.5 * (mx**2 / vx + torch.log(vx))).sum(), mx / vx, 0.5 * (-mx**2 / vx**2 + 1 / vx)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        _check_mode(prop_mode, mc_ok=True)
        N, L, dev, Q = self.N, self.sgp_layer, self.device, self.Din
        sel, lo, cnt, n = self._rows(mb_size)
        eps = _mc_eps(n, Q, dev) if prop_mode == PROP_MC else None
        scale_logZ = -N * 1.0 / n / alpha
        s_cav = -N * 1.0 / n / alpha
        s_post = -N * 1.0 / n * (1.0 - 1.0 / alpha)
        L._fuse_cavity_alpha = alpha
        self.update_hypers(params)
        L.compute_cavity(alpha)
        add = {}
        if cnt > 0:
            yb = self._y[lo:lo + cnt] if sel is None else self._y.index_select(0, sel)
            x1d, x2d = self._x1d, self._x2d
            # cavity of x for this rank's rows (aep_models.py:840-861), one kernel
            mcav, vcav = ops.lvm_x_fwd(0, self.nat_param, x1d, x2d, sel, lo, cnt, self.prior_x1, self.prior_x2, alpha)
            if eps is not None:     # aep_models.py:745-761
                m, v, ctx = L._fwd_mc(mcav, vcav, eps, cav=True)
                dm, dv, logZ, dsn = self.lik_layer._log_Z_mc(m, v, yb, alpha, scale_logZ)
                st = L._bwd_mc(ctx, dm, dv)
            else:
                m, v, ctx = L._fwd_mm(mcav, vcav, cav=True)
                dm, dv, logZ, dsn = self.lik_layer._log_Z(m, v, yb, alpha, scale_logZ)
                st = L._bwd_mm(ctx, dm, dv)
            _add_stats(add, 's_', st)
            # latent-variable terms (aep_models.py:785-801, 817-838, 863-867; base_models.py:913-929): phi_x of
            # cavity and posterior, their chain rules and the layer's input gradients -> x1, x2, one kernel
            add['gx1'], add['gx2'], sums = ops.lvm_x_bwd(0, self.nat_param, x1d, x2d, sel, lo, cnt, self.prior_x1,
                                                         self.prior_x2, alpha, s_cav, s_post,
                                                         st['dmx'].contiguous(), st['dvx'].contiguous())
            add['logZ'], add['dsn'] = logZ.reshape(1), dsn.reshape(1)
            add['phi_cav'], add['phi_post'] = sums[0:1], sums[1:2]
        else:
            _add_stats(add, 's_', _zero_stats(L))
            add['gx1'], add['gx2'] = _zeros(dev, N, Q), _zeros(dev, N, Q)
            for k in ('logZ', 'dsn', 'phi_cav', 'phi_post'):
                add[k] = _zeros(dev, 1)
        # full batch: the x1 / x2 gradients are sharded with the rows -> gathered, not all-reduced
        gx = (add.pop('gx1'), add.pop('gx2')) if sel is None else None
        add = dist.allreduce_dict(add)
        if gx is not None:
            add['gx1'], add['gx2'] = dist.gather_rows(gx[0], N), dist.gather_rows(gx[1], N)
        tail = L._tail_mc if prop_mode == PROP_MC else L._tail_mm
        grads = tail(_get_stats(add, 's_'), alpha)
        if self.lik_layer.has_sn:
            grads['sn'] = add['dsn'].reshape(())
        grads['x1'], grads['x2'] = add['gx1'], add['gx2']
        pm, pv = self.prior_mean, self.prior_var
        phi_prior = 0.5 * (pm**2 / pv + np.log(pv)) * N * Q
        # energy = scale_logZ logZ + [phi_prior + s_cav phi_cav + s_post phi_post] + phi
        energy = tl.dots([(scale_logZ, add['logZ'], None), (s_cav, add['phi_cav'], None),
                          (s_post, add['phi_post'], None), (1.0, L._phi(alpha), None)], const=phi_prior)
        return self._finish(energy, grads, divide_by_N=False)   # aep_models.py:815: no /N


class SGPSSM(Base_SGPSSM):
    """aep_models.py:991-1437."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 x_control=None, gp_emi=False, control_to_emi=True, prec=None, device=None):
        super(SGPSSM, self).__init__(y_train, hidden_size, no_pseudo, lik, prior_mean, prior_var,
                                     x_control, gp_emi, control_to_emi, True, prec, device)
        self.dyn_layer = SGP_Layer(self.N - 1, self.Din + self.Dcon_dyn, self.Din, self.M, True,
                                   prec, self.device)
        if gp_emi:
            self.emi_layer = SGP_Layer(self.N, self.Din + self.Dcon_emi, self.Dout, self.M, True,
                                       prec, self.device)

    def _objective_mm(self, params, alpha, start, end, s_dyn, s_emi):
        """Moment-matched objective with the latent-state algebra in the library's elementwise kernels
        (ops.ssm_*): cavity of every state, tilted transition, the three gradient
        sources chained to the cavity naturals, posterior / cavity log-partitions over all T rows."""
        N, Q, dev = self.N, self.Din, self.device
        dyn, emi = self.dyn_layer, self.emi_layer
        n_emi = end - start
        n_dyn = n_emi - 1
        xf1, xf2 = self._x1d, self._x2d
        pr1, pr2 = self.x_prior_1, self.x_prior_2
        cav_m, cav_v = ops.ssm_cavity(xf1, xf2, pr1, pr2, alpha)
        add = {}
        prev = nxt = up = None
        # ---- transition factors t -> t+1 (aep_models.py:1092-1098, 1317-1348) ---------------
        dlo, dhi = dist.shard(n_dyn)
        t0, t1 = start + dlo, start + dhi
        if t1 > t0:
            mtm1, vtm1 = self._with_control(cav_m[t0:t1], cav_v[t0:t1], t0, t1, self.Dcon_dyn)
            mp, vp, ctx = dyn._fwd_mm(mtm1, vtm1, cav=True)
            dml, dvt, sums = ops.ssm_transition(cav_m[t0 + 1:t1 + 1], cav_v[t0 + 1:t1 + 1], mp, vp, self._sn,
                                                alpha, s_dyn)
            st = dyn._bwd_mm(ctx, dml, dvt)
            sn2 = float(np.exp(2.0 * np.ravel(self.sn)[0]))
            add['logZ_dyn'] = tl.dots([(s_dyn, sums[0:1], None)])
            add['dsn'] = tl.dots([(2.0 * sn2 / alpha, sums[1:2], None)], const=s_dyn * (t1 - t0) * Q * (1 - alpha))
            _add_stats(add, 'd_', st)
            prev = (dml, dvt, t0 + 1)                         # targets of the transitions (dml = -dmt)
            nxt = (st['dmx'], st['dvx'], t0)                  # inputs of the transitions
        else:
            _add_stats(add, 'd_', _zero_stats(dyn))
            add['logZ_dyn'], add['dsn'] = _zeros(dev, 1), _zeros(dev, 1)
        # ---- emission factors (aep_models.py:1100-1113, 1149-1151) --------------------------
        elo, ehi = dist.shard(n_emi)
        e0, e1 = start + elo, start + ehi
        if e1 > e0:
            mup, vup = self._with_control(cav_m[e0:e1], cav_v[e0:e1], e0, e1, self.Dcon_emi)
            yb = self._y[e0:e1]
            if self.gp_emi:
                mo, vo, ctx = emi._fwd_mm(mup, vup, cav=True)
                dme, dve, lZe, dsn_e = self.lik_layer._log_Z(mo, vo, yb, alpha, s_emi)
                ste = emi._bwd_mm(ctx, dme, dve)
                _add_stats(add, 'e_', ste)
                add['logZ_emi'] = tl.dots([(s_emi, lZe.reshape(1), None)])
                add['dsn_emission'] = dsn_e.reshape(1)
                up = (ste['dmx'], ste['dvx'], e0)
            else:
                lZe, dmx, dvx, ge = emi._tilted(mup, vup, alpha, s_emi, yb)
                add['logZ_emi'] = lZe.reshape(1)
                add['dC'], add['dR'] = ge['C'], ge['R']
                up = (dmx, dvx, e0)
        else:
            add['logZ_emi'] = _zeros(dev, 1)
            if self.gp_emi:
                _add_stats(add, 'e_', _zero_stats(emi))
                add['dsn_emission'] = _zeros(dev, 1)
            else:
                add['dC'] = _zeros(dev, self.Dout, Q + self.Dcon_emi)
                add['dR'] = _zeros(dev, self.Dout)
        # the three logZ sources of every latent state -> cavity naturals (aep_models.py:1234-1285)
        add['l1'], add['l2'] = ops.ssm_sources(xf1, xf2, pr1, pr2, alpha, prev, nxt, up)
        add = dist.allreduce_dict(add)

        # ---- replicated tail ----------------------------------------------------------------
        grads = {'sn': add['dsn'].reshape(tuple(np.shape(self.sn)))}
        for k, val in dyn._tail_mm(_get_stats(add, 'd_'), alpha).items():
            grads[k + '_dynamic'] = val
        if self.gp_emi:
            for k, val in emi._tail_mm(_get_stats(add, 'e_'), alpha).items():
                grads[k + '_emission'] = val
            grads['sn_emission'] = add['dsn_emission'].reshape(())
        else:
            grads['C_emission'], grads['R_emission'] = add['dC'], add['dR']
        # x gradients and log-partitions over ALL T rows (aep_models.py:1208-1232, 1287-1315, 1389-1437)
        grads['x_factor_1'], grads['x_factor_2'], sums = ops.ssm_xfinal(xf1, xf2, pr1, pr2, alpha,
                                                                          add['l1'], add['l2'])
        m0, v0 = pr1 / pr2, 1.0 / pr2
        phi_prior = 0.5 * Q * (m0**2 / v0 + np.log(v0))
        terms = [(1.0, add['logZ_dyn'], None), (1.0, add['logZ_emi'], None), (1.0, sums[0:1], None),
                 (1.0, sums[1:2], None), (1.0, dyn._phi(alpha), None)]
        if self.gp_emi:
            terms.append((1.0, emi._phi(alpha), None))
        return self._finish(tl.dots(terms, const=phi_prior), grads)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha=1.0, prop_mode=PROP_MM):
        _check_mode(prop_mode, mc_ok=True)
        N, Q, dev = self.N, self.Din, self.device
        dyn, emi = self.dyn_layer, self.emi_layer
        start, end = self._window(mb_size)
        n_emi = end - start
        n_dyn = n_emi - 1
        mc = prop_mode == PROP_MC
        # Monte-Carlo propagation: the reference draws eps inside the transition forward, then inside
        # the emission forward (aep_models.py:1116, 1129) -- same order here
        eps_dyn = _mc_eps(n_dyn, Q + self.Dcon_dyn, dev) if mc else None
        eps_emi = _mc_eps(n_emi, Q + self.Dcon_emi, dev) if (mc and self.gp_emi) else None
        s_dyn = -(N - 1) * 1.0 / n_dyn / alpha
        s_emi = -N * 1.0 / n_emi / alpha
        dyn._fuse_cavity_alpha = alpha
        if self.gp_emi:
            emi._fuse_cavity_alpha = alpha
        self.update_hypers(params)
        dyn.compute_cavity(alpha)
        if self.gp_emi:
            emi.compute_cavity(alpha)
        if not mc:
            return self._objective_mm(params, alpha, start, end, s_dyn, s_emi)
        # ---- Monte-Carlo propagation: elementwise terms stay on torch (3-D branches) -----------
        # cavity of every latent state (aep_models.py:1376-1387); replicated elementwise work
        f1, f2, p1, p2 = self._f1, self._f2, self._post1, self._post2
        cav1 = p1 - alpha * f1
        cav2 = p2 - alpha * f2
        cav_m, cav_v = cav1 / (cav2 + 1e-16), 1.0 / (cav2 + 1e-16)
        sn2 = torch.exp(2.0 * self._sn)
        add = {'l1': _zeros(dev, N, Q), 'l2': _zeros(dev, N, Q)}

        def push(rows_lo, rows_hi, dmc, dvc):
            """aep_models.py:1252-1275: chain a cavity-moment gradient to the cavity naturals."""
            c1, c2 = cav1[rows_lo:rows_hi], cav2[rows_lo:rows_hi]
            add['l1'][rows_lo:rows_hi] += dmc / c2
            add['l2'][rows_lo:rows_hi] += -dmc * c1 / c2**2 - dvc / c2**2

        # ---- transition factors t -> t+1 (aep_models.py:1092-1098, 1317-1374) ------------
        dlo, dhi = dist.shard(n_dyn)
        t0, t1 = start + dlo, start + dhi
        if t1 > t0:
            mtm1, vtm1 = self._with_control(cav_m[t0:t1], cav_v[t0:t1], t0, t1, self.Dcon_dyn)
            mt, vt = cav_m[t0 + 1:t1 + 1], cav_v[t0 + 1:t1 + 1]
            if mc:
                mp, vp, ctx = dyn._fwd_mc(mtm1.contiguous(), vtm1.contiguous(), eps_dyn, cav=True)
            else:
                mp, vp, ctx = dyn._fwd_mm(mtm1, vtm1, cav=True)
            vsum = vt + vp + sn2 / alpha
            md = mt - mp
            lz = -0.5 * md**2 / vsum - 0.5 * torch.log(1 + alpha * (vt + vp) / sn2) \
                - 0.5 * alpha * torch.log(2 * np.pi * sn2)
            if mc:      # 3-D branch of compute_transition_tilted (aep_models.py:1349-1369)
                lmax = lz.max(dim=0).values
                ex = torch.exp(lz - lmax)
                se = ex.sum(0)
                add['logZ_dyn'] = (s_dyn * (lmax + torch.log(se) - np.log(mp.shape[0])).sum()).reshape(1)
                w = s_dyn * ex / se
                dmp = w * md / vsum
                dvp = w * (-0.5 / vsum + 0.5 * md**2 / vsum**2)
                dmt, dvt = -dmp.sum(0), dvp.sum(0)
                st = dyn._bwd_mc(ctx, dmp, dvp)
            else:
                dvt = s_dyn * (-0.5 / vsum + 0.5 * md**2 / vsum**2)
                dmt = s_dyn * (-md / vsum)
                add['logZ_dyn'] = (s_dyn * lz.sum()).reshape(1)
                st = dyn._bwd_mm(ctx, (-dmt).contiguous(), dvt.contiguous())
            add['dsn'] = (dvt.sum() * 2 * sn2 / alpha + s_dyn * (t1 - t0) * Q * (1 - alpha)).reshape(1)
            _add_stats(add, 'd_', st)
            push(t0 + 1, t1 + 1, dmt, dvt)                                   # "prev" source
            push(t0, t1, st['dmx'][:, :Q], st['dvx'][:, :Q])                 # "next" source
        else:
            _add_stats(add, 'd_', _zero_stats(dyn))
            add['logZ_dyn'], add['dsn'] = _zeros(dev, 1), _zeros(dev, 1)
        # ---- emission factors (aep_models.py:1100-1113, 1149-1151) --------------------------
        elo, ehi = dist.shard(n_emi)
        e0, e1 = start + elo, start + ehi
        if e1 > e0:
            mup, vup = self._with_control(cav_m[e0:e1], cav_v[e0:e1], e0, e1, self.Dcon_emi)
            yb = self._y[e0:e1]
            if self.gp_emi:
                if mc:      # aep_models.py:1127-1145
                    mo, vo, ctx = emi._fwd_mc(mup.contiguous(), vup.contiguous(), eps_emi, cav=True)
                    dme, dve, lZe, dsn_e = self.lik_layer._log_Z_mc(mo, vo, yb, alpha, s_emi)
                    ste = emi._bwd_mc(ctx, dme, dve)
                else:
                    mo, vo, ctx = emi._fwd_mm(mup, vup, cav=True)
                    dme, dve, lZe, dsn_e = self.lik_layer._log_Z(mo, vo, yb, alpha, s_emi)
                    ste = emi._bwd_mm(ctx, dme, dve)
                _add_stats(add, 'e_', ste)
                add['logZ_emi'] = (s_emi * lZe).reshape(1)
                add['dsn_emission'] = dsn_e.reshape(1)
                push(e0, e1, ste['dmx'][:, :Q], ste['dvx'][:, :Q])
            else:
                lZe, dmx, dvx, ge = emi._tilted(mup, vup, alpha, s_emi, yb)
                add['logZ_emi'] = lZe.reshape(1)
                add['dC'], add['dR'] = ge['C'], ge['R']
                push(e0, e1, dmx[:, :Q], dvx[:, :Q])
        else:
            add['logZ_emi'] = _zeros(dev, 1)
            if self.gp_emi:
                _add_stats(add, 'e_', _zero_stats(emi))
                add['dsn_emission'] = _zeros(dev, 1)
            else:
                add['dC'] = _zeros(dev, self.Dout, Q + self.Dcon_emi)
                add['dR'] = _zeros(dev, self.Dout)
        add = dist.allreduce_dict(add)

        # ---- replicated tail ----------------------------------------------------------------
        grads = {'sn': add['dsn'].reshape(tuple(np.shape(self.sn)))}
        for k, val in (dyn._tail_mc if mc else dyn._tail_mm)(_get_stats(add, 'd_'), alpha).items():
            grads[k + '_dynamic'] = val
        if self.gp_emi:
            for k, val in (emi._tail_mc if mc else emi._tail_mm)(_get_stats(add, 'e_'), alpha).items():
                grads[k + '_emission'] = val
            grads['sn_emission'] = add['dsn_emission'].reshape(())
        else:
            grads['C_emission'], grads['R_emission'] = add['dC'], add['dR']
        # x gradients from the posterior / cavity log-partitions over ALL T rows
        # (aep_models.py:1208-1232, 1287-1315) + the logZ sources gathered above (1234-1285)
        one = torch.ones((N, 1), dtype=_F, device=dev)
        s_post = -(1.0 - 1.0 / alpha) * one
        s_post[0:N - 1] += 1.0 / alpha
        s_post[1:N] += 1.0 / alpha
        w3 = 3.0 * one
        w3[0] = 2.0
        w3[-1] = 2.0
        gp1 = s_post * (p1 / p2)
        gp2 = s_post * (-0.5 * p1**2 / p2**2 - 0.5 / p2)
        gx1 = w3 * gp1
        gx2 = 2.0 * w3 * gp2 * f2
        w = w3 - alpha
        sc = (-1.0 / alpha) * one
        sc[0:N - 1] += -1.0 / alpha
        sc[1:N] += -1.0 / alpha
        gx1 = gx1 + (sc * (cav1 / cav2) + add['l1']) * w
        gx2 = gx2 + (sc * (-0.5 * cav1**2 / cav2**2 - 0.5 / cav2) + add['l2']) * w * 2 * f2
        grads['x_factor_1'], grads['x_factor_2'] = gx1, gx2
        # energy (aep_models.py:1186-1197, 1389-1437)
        m0, v0 = self.x_prior_1 / self.x_prior_2, 1.0 / self.x_prior_2
        phi_prior = 0.5 * Q * (m0**2 / v0 + np.log(v0))
        phi_post = (s_post * 0.5 * (p1**2 / p2 - torch.log(p2))).sum()
        phi_cav = (sc * 0.5 * (cav1**2 / cav2 - torch.log(cav2))).sum()
        energy = add['logZ_dyn'] + add['logZ_emi'] + phi_prior + phi_post + phi_cav + dyn._phi(alpha)
        if self.gp_emi:
            energy = energy + emi._phi(alpha)
        return self._finish(energy, grads)
