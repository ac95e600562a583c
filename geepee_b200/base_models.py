"""Model bases: the API surface of the reference (geepee/base_models.py:15-165, 661-1752).

``Base_Model`` keeps the optimiser plumbing on the host exactly as the reference does
(scipy L-BFGS-B / adam over the flattened parameter vector); the model bases upload the
training arrays once, keep them resident in HBM, and own predict_f / predict_y /
init_hypers / get_hypers / update_hypers.
"""
import pickle

import numpy as np
import torch
from scipy.optimize import minimize

from . import config, dist
from . import tail as tl
from .config import PROP_MM, PROP_MC, PROP_LIN
from .layers import _host_copy, default_device, to_dev, pack_to_device, to_host
from .lik_layers import Gauss_Layer, Probit_Layer, Gauss_Emis
from .utils import ObjectiveWrapper, flatten_dict, unflatten_dict, adam, PCA_reduce

_F = torch.float64


def _make_lik(lik, N, Dout, device):
    if lik.lower() == 'gaussian':
        return Gauss_Layer(N, Dout, device)
    if lik.lower() == 'probit':
        return Probit_Layer(N, Dout, device)
    raise NotImplementedError('likelihood not implemented')


class Base_Model(object):
    """base_models.py:15-165."""

    def __init__(self, y_train, prec=None, device=None):
        self.y_train = y_train
        self.N = y_train.shape[0]
        self.fixed_params = []
        self.updated = False
        self.prec = prec
        self.device = device if device is not None else default_device()

    def init_hypers(self, y_train, x_train=None):
        pass

    def get_hypers(self):
        pass

    def update_hypers(self, params):
        pass

    def optimise(self, method='L-BFGS-B', tol=None, reinit_hypers=True, callback=None,
                 maxfun=100000, maxiter=1000, alpha=0.5, mb_size=None, adam_lr=0.001,
                 prop_mode=PROP_MM, disp=True, return_cost=False, **kargs):
        """base_models.py:59-134."""
        self.updated = False
        init_params_dict = self.init_hypers(self.y_train) if reinit_hypers else self.get_hypers()
        init_params_vec, params_args = flatten_dict(init_params_dict)
        objective_wrapper = ObjectiveWrapper()
        if mb_size is None:
            mb_size = self.N
        costs = []
        try:
            if method.lower() == 'adam':
                results = adam(objective_wrapper, init_params_vec, step_size=adam_lr, maxiter=maxiter,
                               args=(params_args, self, mb_size, alpha, prop_mode), disp=disp,
                               callback=callback, return_cost=return_cost)
                if return_cost:
                    final_params, costs = results[0], results[1]
                else:
                    final_params = results
            else:
                options = {'maxfun': maxfun, 'maxiter': maxiter, 'gtol': 1e-6, 'ftol': 1e-6}
                results = minimize(fun=objective_wrapper, x0=init_params_vec,
                                   args=(params_args, self, self.N, alpha, prop_mode), method=method,
                                   jac=True, tol=tol, callback=callback, options=options)
                final_params = results.x
        except KeyboardInterrupt:
            print('Caught KeyboardInterrupt ...')
            final_params = objective_wrapper.previous_x
            costs = []
        final_params = unflatten_dict(final_params, params_args)
        self.update_hypers(final_params)
        if return_cost and method.lower() == 'adam':
            return final_params, costs
        return final_params

    def set_fixed_params(self, params):
        if isinstance(params, (list)):
            for p in params:
                if p not in self.fixed_params:
                    self.fixed_params.append(p)
        else:
            self.fixed_params.append(params)

    def save_model(self, fname='/tmp/model.pickle'):
        pickle.dump(self.get_hypers(), open(fname, 'wb'))

    def load_model(self, fname='/tmp/model.pickle'):
        self.update_hypers(pickle.load(open(fname, 'rb')))

    # ---- helpers shared by the objective functions -----------------------------------------
    def _finish(self, energy, grads, divide_by_N=True):
        """Zero the fixed parameters, divide by N (every model except aep.SGPLVM:
        aep_models.py:663-665 vs 815) and bring energy + all gradients back with ONE
        device->host copy."""
        keys = sorted(grads.keys())
        scale = 1.0 / self.N if divide_by_N else 1.0
        flat = tl.gather([energy.reshape(1)] + [grads[k].reshape(-1) for k in keys], scale)
        host = to_host(flat)
        out, off = {}, 1
        for k in keys:
            n = grads[k].numel()
            arr = np.empty(n, dtype=np.float64)
            _host_copy(arr, host[off:off + n])
            out[k] = arr.reshape(tuple(grads[k].shape))
            off += n
        for p in self.fixed_params:
            out[p] = np.zeros_like(out[p])
        return float(host[0]), out

    def _minibatch_rows(self, mb_size, exact=False):
        """aep_models.py:624-630 (and :712-717 for the LVM's `mb_size == N` test): all rows, or a
        host-side numpy draw from the GLOBAL RNG, as the reference does."""
        N = self.N
        full = (mb_size == N) if exact else (mb_size >= N)
        if full:
            return None
        return dist.agree(np.random.choice(N, mb_size, replace=False))


class Base_SGPR(Base_Model):
    """base_models.py:932-1070."""

    def __init__(self, x_train, y_train, no_pseudo, lik='Gaussian', nat_param=True,
                 prec=None, device=None):
        super(Base_SGPR, self).__init__(y_train, prec, device)
        self.N = y_train.shape[0]
        self.Dout = y_train.shape[1]
        self.Din = x_train.shape[1]
        self.M = no_pseudo
        self.x_train = x_train
        self.nat_param = nat_param
        self.lik_layer = _make_lik(lik, self.N, self.Dout, self.device)
        self._x = to_dev(x_train, self.device)
        self._y = to_dev(y_train, self.device)

    def _batch(self, mb_size):
        """Rows of this call's minibatch owned by this rank -> (xb, yb, batch_size)."""
        idxs = self._minibatch_rows(mb_size)
        n = self.N if idxs is None else idxs.shape[0]
        lo, hi = dist.shard(n)
        if idxs is None:
            return self._x[lo:hi], self._y[lo:hi], n
        sel = torch.as_tensor(idxs[lo:hi], device=self.device)
        return self._x.index_select(0, sel), self._y.index_select(0, sel), n

    def predict_f(self, inputs):
        """base_models.py:985-998."""
        if not self.updated:
            self.sgp_layer.update_posterior()
            self.updated = True
        return self.sgp_layer.forward_prop_thru_post(inputs)

    def sample_f(self, inputs, no_samples=1):
        """base_models.py:1000-1018."""
        if not self.updated:
            self.sgp_layer.update_posterior()
            self.updated = True
        fs = np.zeros((inputs.shape[0], self.Dout, no_samples))
        for k in range(no_samples):
            fs[:, :, k] = self.sgp_layer.sample(inputs)
        return fs

    def predict_y(self, inputs):
        mf, vf = self.predict_f(inputs)
        return self.lik_layer.output_probabilistic(mf, vf)

    def init_hypers(self, y_train):
        init_params = dict(self.sgp_layer.init_hypers(self.x_train))
        init_params.update(self.lik_layer.init_hypers())
        return init_params

    def get_hypers(self):
        params = dict(self.sgp_layer.get_hypers())
        params.update(self.lik_layer.get_hypers())
        return params

    def update_hypers(self, params):
        dev = pack_to_device(params, self.device)
        self.sgp_layer.update_hypers(params, _dev=dev)
        self.lik_layer.update_hypers(params, _dev=dev)


class Base_SDGPR(Base_Model):
    """base_models.py:1073-1336."""

    def __init__(self, x_train, y_train, no_pseudos, hidden_sizes, lik='Gaussian',
                 prec=None, device=None):
        super(Base_SDGPR, self).__init__(y_train, prec, device)
        self.N = y_train.shape[0]
        self.Dout = y_train.shape[1]
        self.Din = x_train.shape[1]
        self.size = [self.Din] + list(hidden_sizes) + [self.Dout]
        self.L = len(self.size) - 1
        if not isinstance(no_pseudos, (list, tuple)):
            self.Ms = [no_pseudos for i in range(self.L)]
        else:
            self.Ms = no_pseudos
        self.x_train = x_train
        self.lik_layer = _make_lik(lik, self.N, self.Dout, self.device)
        self._x = to_dev(x_train, self.device)
        self._y = to_dev(y_train, self.device)

    _batch = Base_SGPR._batch

    def predict_f(self, inputs, prop_mode=PROP_MM, no_samples=200):
        """base_models.py:1132-1158 (moment matching) / 1160-1184 (Monte Carlo)."""
        if prop_mode == PROP_MC:
            return self.predict_f_mc(inputs, no_samples)
        if prop_mode != PROP_MM:
            raise NotImplementedError('prop_mode %s unknown' % prop_mode)
        return self.predict_f_mm(inputs)

    def predict_f_mm(self, inputs):
        """base_models.py:1143-1158."""
        if not self.updated:
            for layer in self.sgp_layers:
                layer.update_posterior()
            self.updated = True
        x = to_dev(inputs, self.device)
        for i, layer in enumerate(self.sgp_layers):
            if i == 0:
                mf, vf, _ = layer._fwd_det(x, cav=False, save=False)
            else:
                mf, vf, _ = layer._fwd_mm(mf, vf, cav=False, save=False)
        return mf.cpu().numpy(), vf.cpu().numpy()

    def predict_f_mc(self, inputs, no_samples):
        """base_models.py:1160-1184: `no_samples` particles per input pushed through the layers
        (deterministic-input kernels on the stacked particles); draws from numpy's global RNG in
        the reference's order.  -> samples[no_samples, n, Dout], and mf, vf of the last layer."""
        if not self.updated:
            for layer in self.sgp_layers:
                layer.update_posterior()
            self.updated = True
        dev = self.device
        samples = to_dev(inputs, dev)
        for i, layer in enumerate(self.sgp_layers):
            mf, vf, _ = layer._fwd_det(samples.contiguous(), cav=False, save=False)
            if i == 0:
                eps = to_dev(np.random.randn(no_samples, mf.shape[0], mf.shape[1]), dev)
                samples = (torch.sqrt(vf) * eps + mf).reshape(no_samples * mf.shape[0], mf.shape[1])
            else:
                eps = to_dev(np.random.randn(mf.shape[0], mf.shape[1]), dev)
                samples = torch.sqrt(vf) * eps + mf
        samples = samples.reshape(no_samples, inputs.shape[0], self.sgp_layers[-1].Dout)
        return samples.cpu().numpy(), mf.cpu().numpy(), vf.cpu().numpy()

    def predict_f_with_input_grad(self, inputs):
        """base_models.py:1186-1237.  As in the reference this is defined for a single GP layer only
        (its deeper branch calls ``backprop_predictive_grads_lvm_mm``, which no layer defines) and
        both gradients it returns are d mf / d inputs (base_models.py:424-425)."""
        if not self.updated:
            for layer in self.sgp_layers:
                layer.update_posterior()
            self.updated = True
        if self.L != 1:
            raise AttributeError("'SGP_Layer' object has no attribute 'backprop_predictive_grads_lvm_mm'")
        x = to_dev(inputs, self.device)
        mf, vf, dx = self.sgp_layers[0]._predictive_dx(x, np.ones((1, 1)), np.zeros((1, 1)))
        dx = dx.cpu().numpy()
        return mf.cpu().numpy(), vf.cpu().numpy(), dx, dx.copy()

    def predict_y_with_input_grad(self, inputs):
        """base_models.py:1277-1289."""
        mf, vf, dm_dx, dv_dx = self.predict_f_with_input_grad(inputs)
        my, vy = self.lik_layer.output_probabilistic(mf, vf)
        return my, vy, dm_dx, dv_dx

    def sample_f(self, inputs, no_samples=1):
        """base_models.py:1239-1262."""
        if not self.updated:
            for layer in self.sgp_layers:
                layer.update_posterior()
            self.updated = True
        fs = np.zeros((inputs.shape[0], self.Dout, no_samples))
        for k in range(no_samples):
            outputs = inputs
            for layer in self.sgp_layers:
                outputs = layer.sample(outputs)
            fs[:, :, k] = outputs
        return fs

    def predict_y(self, inputs):
        mf, vf = self.predict_f(inputs)
        return self.lik_layer.output_probabilistic(mf, vf)

    def init_hypers(self, y_train):
        init_params = dict()
        for i in range(self.L):
            if i == 0:
                sgp_params = self.sgp_layers[i].init_hypers(self.x_train, key_suffix='_%d' % i)
            else:
                sgp_params = self.sgp_layers[i].init_hypers(key_suffix='_%d' % i)
            init_params.update(sgp_params)
        init_params.update(self.lik_layer.init_hypers())
        return init_params

    def get_hypers(self):
        params = dict()
        for i in range(self.L):
            params.update(self.sgp_layers[i].get_hypers(key_suffix='_%d' % i))
        params.update(self.lik_layer.get_hypers())
        return params

    def update_hypers(self, params):
        dev = pack_to_device(params, self.device)
        for i, layer in enumerate(self.sgp_layers):
            layer.update_hypers(params, key_suffix='_%d' % i, _dev=dev)
        self.lik_layer.update_hypers(params, _dev=dev)


class Base_SGPLVM(Base_Model):
    """base_models.py:661-929."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 nat_param=True, prec=None, device=None):
        super(Base_SGPLVM, self).__init__(y_train, prec, device)
        self.N = y_train.shape[0]
        self.Dout = y_train.shape[1]
        self.Din = hidden_size
        self.M = no_pseudo
        self.nat_param = nat_param
        self.lik_layer = _make_lik(lik, self.N, self.Dout, self.device)
        self.factor_x1 = np.zeros((self.N, self.Din))
        self._x2_raw = None         # the log-sqrt parameter as given; factor_x2 = exp(2 x2) on first use
        self._factor_x2 = np.zeros((self.N, self.Din))
        self.prior_mean = prior_mean
        self.prior_var = prior_var
        self.prior_x1 = prior_mean / prior_var
        self.prior_x2 = 1.0 / prior_var
        self._y = to_dev(y_train, self.device)

    def predict_f(self, inputs):
        if not self.updated:
            self.sgp_layer.update_posterior()
            self.updated = True
        return self.sgp_layer.forward_prop_thru_post(inputs)

    def predict_y(self, inputs):
        mf, vf = self.predict_f(inputs)
        return self.lik_layer.output_probabilistic(mf, vf)

    @property
    def factor_x2(self):
        if self._factor_x2 is None:
            self._factor_x2 = np.exp(2 * self._x2_raw)
        return self._factor_x2

    @factor_x2.setter
    def factor_x2(self, value):
        self._factor_x2 = value

    def get_posterior_x(self, idxs=None):
        """base_models.py:765-775."""
        if idxs is None:
            idxs = np.arange(self.N)
        post_1 = self._post1.cpu().numpy()[idxs, :]
        post_2 = self._post2.cpu().numpy()[idxs, :]
        return post_1 / post_2, 1.0 / post_2

    def init_hypers(self, y_train):
        """base_models.py:839-881: PCA latent init + a nested VFE regression fit (which itself
        runs on the device through vfe_models.SGPR)."""
        post_m = PCA_reduce(y_train, self.Din)
        post_m_mean = np.mean(post_m, axis=0)
        post_m_std = np.std(post_m, axis=0) + 1e-5
        post_m = (post_m - post_m_mean) / post_m_std
        post_v = 0.1 * np.ones_like(post_m)
        x_params = {}
        if self.nat_param:
            post_2 = 1.0 / post_v
            x_params['x1'] = post_2 * post_m
            x_params['x2'] = np.log(post_2 - 1) / 2
        else:
            x_params['x1'] = post_m
            x_params['x2'] = np.log(post_v) / 2
        print('init latent function using GPR...')
        from .vfe_models import SGPR
        reg = SGPR(post_m, y_train, self.M, 'Gaussian', self.nat_param, prec=self.prec, device=self.device)
        reg.set_fixed_params(['sn', 'sf', 'ls', 'zu'])
        reg.optimise(method='L-BFGS-B', maxiter=100, disp=False)
        init_params = dict(reg.sgp_layer.get_hypers())
        init_params.update(self.lik_layer.init_hypers())
        init_params.update(x_params)
        return init_params

    def get_hypers(self):
        params = dict(self.sgp_layer.get_hypers())
        params.update(self.lik_layer.get_hypers())
        params['x1'] = self.factor_x1
        params['x2'] = np.log(self.factor_x2) / 2.0
        return params

    def update_hypers(self, params):
        """base_models.py:899-911."""
        dev = pack_to_device(params, self.device)
        self.sgp_layer.update_hypers(params, _dev=dev)
        self.lik_layer.update_hypers(params, _dev=dev)
        self.factor_x1 = params['x1']
        # the reference stores exp(2 x2) here (base_models.py:903); only get_hypers reads it back, and N Q
        # host exponentials per objective call cost milliseconds at N = 1e5 .. 1e6 -> formed on first use
        self._x2_raw, self._factor_x2 = params['x2'], None
        # raw device parameters: the objective's latent-variable kernels (ops.lvm_x_fwd / lvm_x_bwd) read
        # them directly; the derived arrays below exist for the prediction / inspection API only
        self._x1d = dev['x1'].reshape(self.N, self.Din)
        self._x2d = dev['x2'].reshape(self.N, self.Din)
        self._lazy = {}

    @property
    def _f1(self):
        return self._x1d

    @property
    def _f2(self):
        if 'f2' not in self._lazy:
            self._lazy['f2'] = torch.exp(2.0 * self._x2d)
        return self._lazy['f2']

    @property
    def _post1(self):
        if 'p1' not in self._lazy:
            self._lazy['p1'] = (self.prior_x1 + self._f1) if self.nat_param else self._f1 / self._f2
        return self._lazy['p1']

    @property
    def _post2(self):
        if 'p2' not in self._lazy:
            self._lazy['p2'] = (self.prior_x2 + self._f2) if self.nat_param else 1.0 / self._f2
        return self._lazy['p2']

    def _rows(self, mb_size):
        """This rank's rows of the minibatch -> (sel, lo, cnt, n): `sel` = device int64 indices of a drawn
        minibatch (cnt of them), or None for the contiguous rows lo .. lo+cnt-1 of the full batch; n = size of
        the whole minibatch."""
        idxs = self._minibatch_rows(mb_size, exact=True)
        n = self.N if idxs is None else idxs.shape[0]
        lo, hi = dist.shard(n)
        if idxs is None:
            return None, lo, hi - lo, n
        return torch.as_tensor(idxs[lo:hi], device=self.device), 0, hi - lo, n


class Base_SGPSSM(Base_Model):
    """base_models.py:1339-1752."""

    def __init__(self, y_train, hidden_size, no_pseudo, lik='Gaussian', prior_mean=0, prior_var=1,
                 x_control=None, gp_emi=False, control_to_emi=True, nat_param=True,
                 prec=None, device=None):
        super(Base_SGPSSM, self).__init__(y_train, prec, device)
        if x_control is not None:
            self.Dcon_dyn = x_control.shape[1]
            self.x_control = x_control
            self.Dcon_emi = x_control.shape[1] if control_to_emi else 0
            self._xc = to_dev(x_control, self.device)
        else:
            self.Dcon_dyn = 0
            self.Dcon_emi = 0
            self.x_control = None
            self._xc = None
        self.N = y_train.shape[0]
        self.Dout = y_train.shape[1]
        self.Din = hidden_size
        self.M = no_pseudo
        self.gp_emi = gp_emi
        self.nat_param = nat_param
        if gp_emi:
            self.lik_layer = _make_lik(lik, self.N, self.Dout, self.device)
        else:
            if lik.lower() != 'gaussian':
                raise NotImplementedError('likelihood not implemented')
            self.emi_layer = Gauss_Emis(y_train, self.Dout, self.Din + self.Dcon_emi, self.device)
        self.prior_mean = prior_mean
        self.prior_var = prior_var
        self.x_prior_1 = prior_mean / prior_var
        self.x_prior_2 = 1.0 / prior_var
        self._y = to_dev(y_train, self.device)

    def predict_f(self, inputs):
        return self.dyn_layer.forward_prop_thru_post(inputs)

    def predict_y(self, inputs):
        """base_models.py:1544-1560: one transition step from given states, then the emission."""
        mf, vf = self.dyn_layer.forward_prop_thru_post(inputs)
        if self.gp_emi:
            mg, vg = self.emi_layer.forward_prop_thru_post(mf, vf)
            return self.lik_layer.output_probabilistic(mg, vg)
        my, _, vy = self.emi_layer.output_probabilistic(mf, vf)
        return my, np.diagonal(vy, axis1=1, axis2=2)

    def predict_forward(self, T, x_control=None, prop_mode=PROP_MM, no_samples=200):
        """base_models.py:1453-1461."""
        if prop_mode == PROP_MM:
            return self.predict_forward_mm(T, x_control)
        if prop_mode == PROP_LIN:
            raise NotImplementedError('TODO')
        if prop_mode == PROP_MC:
            return self.predict_forward_mc(T, x_control, no_samples)
        raise NotImplementedError('unknown prop mode %s' % prop_mode)

    def predict_forward_mm(self, T, x_control):
        """base_models.py:1463-1502: T-step roll-out from the last posterior state with moment
        matching (a sequential chain of one-row layer evaluations: launch bound by nature)."""
        mx, vx = np.zeros((T, self.Din)), np.zeros((T, self.Din))
        my, vy_noiseless, vy = (np.zeros((T, self.Dout)) for _ in range(3))
        post_m, post_v = self.get_posterior_x()
        mtm1, vtm1 = post_m[[-1], :], post_v[[-1], :]
        for t in range(T):
            if self.Dcon_dyn > 0:
                mtm1 = np.hstack((mtm1, x_control[[t], :]))
                vtm1 = np.hstack((vtm1, np.zeros((1, self.Dcon_dyn))))
            mt, vt = self.dyn_layer.forward_prop_thru_post(mtm1, vtm1)
            if self.Dcon_emi > 0:
                mtc = np.hstack((mt, x_control[[t], :]))
                vtc = np.hstack((vt, np.zeros((1, self.Dcon_emi))))
            else:
                mtc, vtc = mt, vt
            if self.gp_emi:
                mft, vft = self.emi_layer.forward_prop_thru_post(mtc, vtc)
                myt, vyt_n = self.lik_layer.output_probabilistic(mft, vft)
            else:
                # (the reference passes the un-augmented state here, base_models.py:1492, which
                #  only works without control inputs to the emission; the augmented one is meant)
                myt, vyt, vyt_n = self.emi_layer.output_probabilistic(mtc, vtc)
                vft = np.diagonal(vyt, axis1=1, axis2=2)
                vyt_n = np.diagonal(vyt_n, axis1=1, axis2=2)
            mx[t, :], vx[t, :] = mt, vt
            my[t, :], vy_noiseless[t, :], vy[t, :] = myt, vft, vyt_n
            mtm1, vtm1 = mt, vt
        return mx, vx, my, vy_noiseless, vy

    def predict_forward_mc(self, T, x_control, no_samples):
        """base_models.py:1504-1542: roll-out of `no_samples` particles; the standard-normal draws
        come from numpy's global RNG in the reference's order."""
        x = np.zeros((T, no_samples, self.Din))
        my, vy = np.zeros((T, no_samples, self.Dout)), np.zeros((T, no_samples, self.Dout))
        post_m, post_v = self.get_posterior_x()
        mtm1, vtm1 = post_m[[-1], :], post_v[[-1], :]
        eps = np.random.randn(no_samples, self.Din)
        x_samples = eps * np.sqrt(vtm1) + mtm1
        for t in range(T):
            if self.Dcon_dyn > 0:
                xc_samples = np.hstack((x_samples, np.tile(x_control[[t], :], [no_samples, 1])))
            else:
                xc_samples = x_samples
            mt, vt = self.dyn_layer.forward_prop_thru_post(xc_samples)
            eps = np.random.randn(no_samples, self.Din)
            x_samples = eps * np.sqrt(vt) + mt
            if self.Dcon_emi > 0:
                xc_samples = np.hstack((x_samples, np.tile(x_control[[t], :], [no_samples, 1])))
            else:
                xc_samples = x_samples
            if self.gp_emi:
                mft, vft = self.emi_layer.forward_prop_thru_post(xc_samples)
                myt, vyt_n = self.lik_layer.output_probabilistic(mft, vft)
            else:
                myt, _, vyt_n = self.emi_layer.output_probabilistic(xc_samples, np.zeros_like(xc_samples))
                vyt_n = np.diagonal(vyt_n, axis1=1, axis2=2)
            x[t, :, :] = x_samples
            my[t, :, :], vy[t, :, :] = myt, vyt_n
        return x, my, vy

    def get_posterior_x(self, idxs=None):
        p1, p2 = self._post1.cpu().numpy(), self._post2.cpu().numpy()
        if idxs is not None:
            p1, p2 = p1[idxs, :], p2[idxs, :]
        return p1 / p2, 1.0 / p2

    def get_posterior_y(self):
        """base_models.py:1578-1595."""
        mx, vx = self.get_posterior_x()
        if self.Dcon_emi > 0:
            mx = np.hstack((mx, self.x_control))
            vx = np.hstack((vx, np.zeros((self.N, self.Dcon_emi))))
        if self.gp_emi:
            mf, vf = self.emi_layer.forward_prop_thru_post(mx, vx)
            my, vyn = self.lik_layer.output_probabilistic(mf, vf)
        else:
            my, vy, vyn = self.emi_layer.output_probabilistic(mx, vx)
            vf = np.diagonal(vy, axis1=1, axis2=2)
            vyn = np.diagonal(vyn, axis1=1, axis2=2)
        return my, vf, vyn

    @property
    def x_factor_2(self):
        if self._x_factor_2 is None:
            self._x_factor_2 = np.exp(2 * self._xf2_raw)
        return self._x_factor_2

    @x_factor_2.setter
    def x_factor_2(self, value):
        self._x_factor_2 = value

    def get_hypers(self):
        params = dict(self.dyn_layer.get_hypers(key_suffix='_dynamic'))
        params.update(self.emi_layer.get_hypers(key_suffix='_emission'))
        params['x_factor_1'] = self.x_factor_1
        params['x_factor_2'] = np.log(self.x_factor_2) / 2.0
        params['sn'] = self.sn
        if self.gp_emi:
            params.update(self.lik_layer.get_hypers(key_suffix='_emission'))
        return params

    def update_hypers(self, params):
        """base_models.py:1711-1728."""
        dev = self.device
        dp = pack_to_device(params, dev)
        self.dyn_layer.update_hypers(params, key_suffix='_dynamic', _dev=dp)
        self.emi_layer.update_hypers(params, key_suffix='_emission', _dev=dp)
        if self.gp_emi:
            self.lik_layer.update_hypers(params, key_suffix='_emission', _dev=dp)
        self.sn = params['sn']
        self._sn = dp['sn'].reshape(-1)[:1].contiguous()
        self.x_factor_1 = params['x_factor_1']
        # exp(2 x_factor_2) (base_models.py:1716) is only read back by get_hypers: formed on first use
        self._xf2_raw, self._x_factor_2 = params['x_factor_2'], None
        # raw device parameters: the moment-matched AEP objective's latent-state kernels (ops.ssm_*) read
        # them directly; the derived arrays below are formed on first use (Monte-Carlo / VFE / prediction paths)
        self._x1d = dp['x_factor_1'].reshape(self.N, self.Din)
        self._x2d = dp['x_factor_2'].reshape(self.N, self.Din)
        self._lazy = {}

    @property
    def _f1(self):
        return self._x1d

    @property
    def _f2(self):
        if 'f2' not in self._lazy:
            self._lazy['f2'] = torch.exp(2.0 * self._x2d)
        return self._lazy['f2']

    def _posts(self):
        if 'p1' not in self._lazy:
            if self.nat_param:      # base_models.py:1719-1725: factors tied 3x (2x at the ends) + prior at t = 0
                w = torch.full((self.N, 1), 3.0, dtype=_F, device=self.device)
                w[0] = 2.0
                w[-1] = 2.0
                p1 = w * self._f1
                p2 = w * self._f2
                p1[0] += self.x_prior_1
                p2[0] += self.x_prior_2
            else:
                p1 = self._f1 / self._f2
                p2 = 1.0 / self._f2
            self._lazy['p1'], self._lazy['p2'] = p1, p2
        return self._lazy['p1'], self._lazy['p2']

    @property
    def _post1(self):
        return self._posts()[0]

    @property
    def _post2(self):
        return self._posts()[1]

    def _window(self, mb_size):
        """aep_models.py:1045-1057: whole series or one random contiguous window."""
        N = self.N
        if mb_size >= N:
            return 0, N
        start = int(dist.agree(np.random.randint(0, N - mb_size)))
        return start, start + mb_size

    def _with_control(self, m, v, lo, hi, Dcon):
        if Dcon > 0:
            xc = self._xc[lo:hi]
            return (torch.cat((m, xc), dim=1).contiguous(),
                    torch.cat((v, torch.zeros_like(xc)), dim=1).contiguous())
        return m.contiguous(), v.contiguous()

    def init_hypers(self, y_train):
        """base_models.py:1605-1693 for Din == Dout (the LDS initialiser needs pylds)."""
        if self.Din != self.Dout:
            raise NotImplementedError('init_hypers with Din != Dout needs pylds (base_models.py:1611)')
        post_m = np.copy(y_train)
        post_v = 0.1 * np.ones_like(post_m)
        ssm_params = {'sn': np.log(0.01) * np.ones(1)}
        if self.nat_param:
            post_2 = 1.0 / post_v
            ssm_params['x_factor_1'] = post_2 * post_m / 3
            ssm_params['x_factor_2'] = np.log(post_2 / 3) / 2
        else:
            ssm_params['x_factor_1'] = np.copy(post_m)
            ssm_params['x_factor_2'] = np.log(post_v) / 2
        print('init latent function using GPR...')
        x = post_m[:self.N - 1, :]
        y = post_m[1:, :]
        if self.Dcon_dyn > 0:
            x = np.hstack((x, self.x_control[:self.N - 1, :]))
        from .vfe_models import SGPR
        reg = SGPR(x, y, self.M, 'Gaussian', self.nat_param, prec=self.prec, device=self.device)
        reg.set_fixed_params(['sn', 'sf'])
        opt_params = reg.optimise(method='L-BFGS-B', maxiter=500, disp=False)
        reg.update_hypers(opt_params)
        init_params = dict(reg.sgp_layer.get_hypers(key_suffix='_dynamic'))
        if self.gp_emi:
            print('init emission function using GPR...')
            x = post_m
            if self.Dcon_emi > 0:
                x = np.hstack((x, self.x_control))
            reg = SGPR(x, self.y_train, self.M, 'Gaussian', self.nat_param, prec=self.prec, device=self.device)
            reg.set_fixed_params(['sn', 'sf', 'ls', 'zu'])
            opt_params = reg.optimise(method='L-BFGS-B', alpha=0.5, maxiter=5000, disp=False)
            reg.update_hypers(opt_params)
            init_params.update(reg.sgp_layer.get_hypers(key_suffix='_emission'))
            init_params.update(self.lik_layer.init_hypers(key_suffix='_emission'))
        else:
            emi_params = self.emi_layer.init_hypers(key_suffix='_emission')
            emi_params['C_emission'] = np.eye(self.Din)
            init_params.update(emi_params)
        init_params.update(ssm_params)
        return init_params
