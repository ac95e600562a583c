"""VFE (uncollapsed variational free energy) models on the B200 hot path.

Reference: geepee/vfe_models.py -- SGP_Layer 290-548, SGPR 551-719, SGPLVM 722-863,
SGPSSM 866-1119.  Same kernels as the AEP path with posterior (A, B) operands instead of
cavity ones and a KL tail instead of the log-partition tail.  ``SGPR_collapsed`` (15-289) is
a different, full-batch algorithm and out of scope.
"""
import numpy as np
import torch

from . import dist, nvtx, ops
from . import tail as tl
from .aep_models import _add_stats, _get_stats, _zero_stats, _zeros, _check_mode, _mc_eps
from .base_models import Base_SGPR, Base_SGPLVM, Base_SGPSSM
from .config import PROP_MM, PROP_MC
from .layers import VFE_SGP_Layer as SGP_Layer  # noqa: F401  (reference name)

_F = torch.float64


class SGPR(Base_SGPR):
    """vfe_models.py:551-632."""

    def __init__(self, x_train, y_train, no_pseudo, lik='Gaussian', nat_param=True,
                 prec=None, device=None):
        super(SGPR, self).__init__(x_train, y_train, no_pseudo, lik, nat_param, prec, device)
        self.sgp_layer = SGP_Layer(self.N, self.Din, self.Dout, self.M, nat_param, prec, self.device)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha='not_used', prop_mode='not_used'):
        N, L, dev = self.N, self.sgp_layer, self.device
        xb, yb, n = self._batch(mb_size)
        scale = -N * 1.0 / n
        self.update_hypers(params)
        add = {}
        if xb.shape[0] > 0:
            def lik(m, v, c0, c1):
                dm, dv, ll, dsn = self.lik_layer._log_lik_exp(m, v, yb[c0:c1], scale)
                return dm, dv, {'ll': ll.reshape(1), 'dsn': dsn.reshape(1)}
            st, ext = L.det_step(xb, lik, cav=False)      # row-chunked: saved Kfu / T stay bounded
            _add_stats(add, 's_', st)
            add.update(ext)
        else:
            _add_stats(add, 's_', _zero_stats(L))
            add['ll'], add['dsn'] = _zeros(dev, 1), _zeros(dev, 1)
        add = dist.allreduce_dict(add)
        grads = L._tail(_get_stats(add, 's_'), False)
        if self.lik_layer.has_sn:
            grads['sn'] = add['dsn'].reshape(())
        energy = tl.dots([(scale, add['ll'], None), (1.0, L._kl(), None)])
        return self._finish(energy, grads)


class SGPLVM(Base_SGPLVM):
    """vfe_models.py:722-863."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 nat_param=True, prec=None, device=None):
        super(SGPLVM, self).__init__(y_train, hidden_size, no_pseudo, lik, prior_mean, prior_var,
                                     nat_param, prec, device)
        self.sgp_layer = SGP_Layer(self.N, self.Din, self.Dout, self.M, nat_param, prec, self.device)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha='not_used', prop_mode=PROP_MM):
        _check_mode(prop_mode, mc_ok=True)
        N, L, dev, Q = self.N, self.sgp_layer, self.device, self.Din
        sel, lo, cnt, n = self._rows(mb_size)
        eps = _mc_eps(n, Q, dev) if prop_mode == PROP_MC else None
        scale = -N * 1.0 / n
        sx = N * 1.0 / n
        self.update_hypers(params)
        m0, v0 = self.prior_mean, self.prior_var
        add = {}
        if cnt > 0:
            yb = self._y[lo:lo + cnt] if sel is None else self._y.index_select(0, sel)
            x1d, x2d = self._x1d, self._x2d
            # posterior of x for this rank's rows (base_models.py:765-775), one kernel
            mx, vx = ops.lvm_x_fwd(1, self.nat_param, x1d, x2d, sel, lo, cnt, m0, v0, 1.0)
            if eps is not None:     # vfe_models.py:793-808: sample average of the expected log-lik
                K = eps.shape[0]
                m, v, ctx = L._fwd_mc(mx, vx, eps, cav=False)
                dm, dv, ll, dsn = self.lik_layer._log_lik_exp(
                    m.reshape(-1, self.Dout), v.reshape(-1, self.Dout), yb.repeat(K, 1), scale / K)
                ll = ll / K
                st = L._bwd_mc(ctx, dm, dv)
            else:
                m, v, ctx = L._fwd_mm(mx, vx, cav=False)
                dm, dv, ll, dsn = self.lik_layer._log_lik_exp(m, v, yb, scale)
                st = L._bwd_mm(ctx, dm, dv)
            _add_stats(add, 's_', st)
            # KL of q(x) (vfe_models.py:857-863) and the chain to x1, x2 (base_models.py:913-929), one kernel
            add['gx1'], add['gx2'], sums = ops.lvm_x_bwd(1, self.nat_param, x1d, x2d, sel, lo, cnt, m0, v0, 1.0, sx, 0.0,
                                                         st['dmx'].contiguous(), st['dvx'].contiguous())
            add['ll'], add['dsn'], add['klx'] = ll.reshape(1), dsn.reshape(1), sums[0:1]
        else:
            _add_stats(add, 's_', _zero_stats(L))
            add['gx1'], add['gx2'] = _zeros(dev, N, Q), _zeros(dev, N, Q)
            for k in ('ll', 'dsn', 'klx'):
                add[k] = _zeros(dev, 1)
        # full batch: the x1 / x2 gradients are sharded with the rows -> gathered, not all-reduced
        gx = (add.pop('gx1'), add.pop('gx2')) if sel is None else None
        add = dist.allreduce_dict(add)
        if gx is not None:
            add['gx1'], add['gx2'] = dist.gather_rows(gx[0], N), dist.gather_rows(gx[1], N)
        grads = L._tail(_get_stats(add, 's_'), prop_mode != PROP_MC)
        if self.lik_layer.has_sn:
            grads['sn'] = add['dsn'].reshape(())
        grads['x1'], grads['x2'] = add['gx1'], add['gx2']
        energy = tl.dots([(scale, add['ll'], None), (sx, add['klx'], None), (1.0, L._kl(), None)])
        return self._finish(energy, grads)


class SGPSSM(Base_SGPSSM):
    """vfe_models.py:866-1119."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 x_control=None, gp_emi=False, control_to_emi=True, nat_param=True,
                 prec=None, device=None):
        super(SGPSSM, self).__init__(y_train, hidden_size, no_pseudo, lik, prior_mean, prior_var,
                                     x_control, gp_emi, control_to_emi, nat_param, prec, device)
        self.dyn_layer = SGP_Layer(self.N - 1, self.Din + self.Dcon_dyn, self.Din, self.M, nat_param,
                                   prec, self.device)
        if gp_emi:
            self.emi_layer = SGP_Layer(self.N, self.Din + self.Dcon_emi, self.Dout, self.M, nat_param,
                                       prec, self.device)

    @nvtx.annotate('objective_function')
    def objective_function(self, params, mb_size, alpha='not_used', prop_mode=PROP_MM):
        _check_mode(prop_mode, mc_ok=True)
        N, Q, dev = self.N, self.Din, self.device
        dyn, emi = self.dyn_layer, self.emi_layer
        start, end = self._window(mb_size)
        nb = end - start
        mc = prop_mode == PROP_MC       # eps in the reference's draw order (vfe_models.py:985, 998)
        eps_dyn = _mc_eps(nb - 1, Q + self.Dcon_dyn, dev) if mc else None
        eps_emi = _mc_eps(nb, Q + self.Dcon_emi, dev) if (mc and self.gp_emi) else None
        s_dyn = -(N - 1) * 1.0 / (nb - 1)
        s_emi = -N * 1.0 / nb
        s_ent = -N * 1.0 / nb
        self.update_hypers(params)
        p1, p2 = self._post1, self._post2
        pm, pv = p1 / p2, 1.0 / p2          # posterior of every latent state (elementwise)
        sn2 = torch.exp(2.0 * self._sn)
        add = {'dm': _zeros(dev, N, Q), 'dv': _zeros(dev, N, Q)}
        # ---- transitions (vfe_models.py:941-948, 1080-1092) ---------------------------------
        dlo, dhi = dist.shard(nb - 1)
        t0, t1 = start + dlo, start + dhi
        if t1 > t0:
            mtm1, vtm1 = self._with_control(pm[t0:t1], pv[t0:t1], t0, t1, self.Dcon_dyn)
            mt, vt = pm[t0 + 1:t1 + 1], pv[t0 + 1:t1 + 1]
            if mc:
                mp, vp, ctx = dyn._fwd_mc(mtm1.contiguous(), vtm1.contiguous(), eps_dyn, cav=False)
            else:
                mp, vp, ctx = dyn._fwd_mm(mtm1, vtm1, cav=False)
            K = mp.shape[0] if mc else 1        # 3-D branch: sample average (vfe_models.py:1094-1104)
            t2 = -0.5 / sn2 * (mt**2 + vt - 2 * mt * mp + mp**2 + vp)
            add['logZ_dyn'] = (s_dyn * (-0.5 * torch.log(2 * np.pi * sn2) * t2.numel() + t2.sum()) / K).reshape(1)
            add['dsn'] = (s_dyn * (-1 - 2 * t2).sum() / K).reshape(1)
            if mc:
                dmp = s_dyn / sn2 * (mt - mp) / K
                dvp = -s_dyn * 0.5 / sn2 * torch.ones_like(vp) / K
                dmt, dvt = -dmp.sum(0), dvp.sum(0)
                st = dyn._bwd_mc(ctx, dmp, dvp)
            else:
                dmt = -s_dyn / sn2 * (mt - mp)
                dvt = -s_dyn * 0.5 / sn2 * torch.ones_like(vt)
                st = dyn._bwd_mm(ctx, (-dmt).contiguous(), dvt.contiguous())
            _add_stats(add, 'd_', st)
            add['dm'][t0 + 1:t1 + 1] += dmt
            add['dv'][t0 + 1:t1 + 1] += dvt
            add['dm'][t0:t1] += st['dmx'][:, :Q]
            add['dv'][t0:t1] += st['dvx'][:, :Q]
        else:
            _add_stats(add, 'd_', _zero_stats(dyn))
            add['logZ_dyn'], add['dsn'] = _zeros(dev, 1), _zeros(dev, 1)
        # ---- emissions ------------------------------------------------------------------------
        elo, ehi = dist.shard(nb)
        e0, e1 = start + elo, start + ehi
        if e1 > e0:
            mup, vup = self._with_control(pm[e0:e1], pv[e0:e1], e0, e1, self.Dcon_emi)
            yb = self._y[e0:e1]
            if self.gp_emi and mc:      # vfe_models.py:996-1010
                K = eps_emi.shape[0]
                mo, vo, ctx = emi._fwd_mc(mup.contiguous(), vup.contiguous(), eps_emi, cav=False)
                dme, dve, lle, dsn_e = self.lik_layer._log_lik_exp(
                    mo.reshape(-1, self.Dout), vo.reshape(-1, self.Dout), yb.repeat(K, 1), s_emi / K)
                lle = lle / K
                ste = emi._bwd_mc(ctx, dme, dve)
                _add_stats(add, 'e_', ste)
                add['logZ_emi'] = (s_emi * lle).reshape(1)
                add['dsn_emission'] = dsn_e.reshape(1)
                dmx, dvx = ste['dmx'], ste['dvx']
            elif self.gp_emi:
                mo, vo, ctx = emi._fwd_mm(mup, vup, cav=False)
                dme, dve, lle, dsn_e = self.lik_layer._log_lik_exp(mo, vo, yb, s_emi)
                ste = emi._bwd_mm(ctx, dme, dve)
                _add_stats(add, 'e_', ste)
                add['logZ_emi'] = (s_emi * lle).reshape(1)
                add['dsn_emission'] = dsn_e.reshape(1)
                dmx, dvx = ste['dmx'], ste['dvx']
            else:
                lZe, dmx, dvx, ge = emi._log_lik_exp(mup, vup, s_emi, yb)
                add['logZ_emi'] = lZe.reshape(1)
                add['dC'], add['dR'] = ge['C'], ge['R']
            add['dm'][e0:e1] += dmx[:, :Q]
            add['dv'][e0:e1] += dvx[:, :Q] + s_ent * 0.5 / pv[e0:e1]       # entropy term
            add['ent'] = (0.5 * torch.log(pv[e0:e1])).sum().reshape(1)
        else:
            add['logZ_emi'], add['ent'] = _zeros(dev, 1), _zeros(dev, 1)
            if self.gp_emi:
                _add_stats(add, 'e_', _zero_stats(emi))
                add['dsn_emission'] = _zeros(dev, 1)
            else:
                add['dC'] = _zeros(dev, self.Dout, Q + self.Dcon_emi)
                add['dR'] = _zeros(dev, self.Dout)
        add = dist.allreduce_dict(add)

        grads = {'sn': add['dsn'].reshape(tuple(np.shape(self.sn)))}
        for k, val in dyn._tail(_get_stats(add, 'd_'), not mc).items():
            grads[k + '_dynamic'] = val
        if self.gp_emi:
            for k, val in emi._tail(_get_stats(add, 'e_'), not mc).items():
                grads[k + '_emission'] = val
            grads['sn_emission'] = add['dsn_emission'].reshape(())
        else:
            grads['C_emission'], grads['R_emission'] = add['dC'], add['dR']
        # base_models.py:1730-1752 compute_posterior_grad_x (rows outside the window stay 0)
        dm, dv = add['dm'], add['dv']
        f2 = self._f2
        if self.nat_param:
            w = torch.full((N, 1), 3.0, dtype=_F, device=dev)
            w[0] = 2.0
            w[-1] = 2.0
            g1 = dm / p2 * w
            g2 = (-dm * p1 / p2**2 - dv / p2**2) * w * 2 * f2
        else:
            g1 = dm
            g2 = dv * 2 * f2
        grads['x_factor_1'], grads['x_factor_2'] = g1, g2
        x_ent = s_ent * (nb * Q * (0.5 + 0.5 * np.log(2 * np.pi)) + add['ent'])
        energy = add['logZ_dyn'] + add['logZ_emi'] + x_ent + dyn._kl()
        if self.gp_emi:
            energy = energy + emi._kl()
        return self._finish(energy, grads)
