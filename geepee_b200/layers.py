"""Sparse-GP layer: host-side mirror of the reference's SGP_Layer family.

Reference: geepee/base_models.py:168-658 (Base_SGP_Layer), geepee/aep_models.py:26-586
(AEP SGP_Layer), geepee/vfe_models.py:290-548 (VFE SGP_Layer).  Same method names, argument
meaning and parameter dict keys.

Division of labour
  * every O(n) pass over rows runs in the CUDA kernels of libgeepee_b200.so (ops.py):
    Kfu / psi generation fused with the contractions; the kernels return only the
    cross-row sufficient statistics  S = {dA, dB, dzu, dl, dsf2, ...}  (SURVEY.md section 8a);
  * the data-independent O(Dout M^3) "tail" (q(u) algebra, cavity, log-partitions, chain
    rules back to eta1_R / eta2 / kernel hypers) is evaluated here in fp64 on the device with
    batched Cholesky + GEMM calls; it is identical for every rank of a data-parallel job.
All state lives in device tensors; the numpy views the reference exposes (``layer.Kuu`` ...)
are materialised on attribute access for the layer-level API and the tests.
"""
import numpy as np
import torch

from . import config, nvtx, ops, _lib
from . import tail as tl
from .tailgraph import TailGraph

_F = torch.float64

# the cross-row statistics a post-tail consumes
_TAIL_STAT_KEYS = ('dA', 'dB', 'dzu', 'dl', 'dsf2', 'dvsum')

# device tensors exposed under the reference's attribute names
_EXPORTED = ('Kuu', 'Kuuinv', 'Su', 'Suinv', 'mu', 'Splusmm', 'A', 'B_det', 'B_sto', 'theta_1',
             'theta_1_R', 'theta_2', 'Suhat', 'Suhatinv', 'muhat', 'Splusmmhat', 'Ahat', 'Bhat_det',
             'Bhat_sto')


def default_device():
    if _lib.device_type() == 'cuda':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')       # only reachable through the test hook (emulator)


def to_dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device)


_PIN = {}   # pinned staging buffers, one per device (grown on demand)


def _pinned(device, n, slot):
    key = (str(device), slot)
    buf = _PIN.get(key)
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1024), dtype=torch.float64, pin_memory=True)
        _PIN[key] = buf
    return buf


_ZC_MAX = 1 << 20       # elements: uploads up to this size are read from pinned memory by a copy kernel
_ZC = {}                # pinned buffer address -> device-side address (0: not mapped, use the copy engine)


def _zero_copy_upload(pin, t, total):
    """t[:total] = pin[:total] by a copy kernel that reads the pinned staging buffer over PCIe (gpb_tail_copy), so that
    the per-step parameter upload does not queue on the H2D copy engine behind a large input copy in flight.
    GPB_UPLOAD_DMA=1 forces the copy engine.  Returns False when the pinned buffer is not device-mapped."""
    import ctypes
    import os
    if os.environ.get('GPB_UPLOAD_DMA'):
        return False
    from . import _lib as _l, ops
    lib = _l.get()
    hp = pin.data_ptr()
    dp = _ZC.get(hp)
    if dp is None:
        out = ctypes.c_void_p()
        rc = lib.gpb_host_device_ptr(ctypes.c_void_p(hp), ctypes.byref(out))
        dp = out.value if rc == 0 and out.value else 0
        _ZC[hp] = dp
    if not dp:
        return False
    srcs = (ctypes.c_void_p * 1)(dp)
    dsts = (ctypes.c_void_p * 1)(t.data_ptr())
    counts = (ctypes.c_long * 1)(total)
    ops._chk(lib.gpb_tail_copy(1, srcs, dsts, counts, ops._stream(t)), 'tail_copy')
    return True


_BIG = 1 << 20          # elements: arrays above this are copied by several host threads
_POOL = None


def _host_copy(dst, src):
    """dst[:] = src for flat fp64 numpy views; large copies are split over a few threads (numpy
    releases the GIL in memcpy; one core moves ~8 GB/s, the latent arrays of an SGPSSM / SGPLVM with
    1e6 rows are 32 MB per parameter and direction)."""
    n = src.size
    if n < _BIG:
        dst[:] = src
        return
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=4)
    step = (n + 3) // 4

    def part(i):
        dst[i:i + step] = src[i:i + step]
    list(_POOL.map(part, range(0, n, step)))


def pack_to_device(params, device):
    """Upload a whole parameter dict with ONE host->device copy; returns {key: fp64 device view}.
    On a GPU the arrays are gathered straight into a pinned staging buffer (no pageable bounce:
    the latent arrays of SGPLVM / SGPSSM are [N,Q] parameters, 64 MB at T = 1e6)."""
    keys = sorted(params.keys())
    arrs = [np.asarray(params[k], dtype=np.float64) for k in keys]
    total = int(sum(a.size for a in arrs))
    if device.type == 'cuda' and total > 0:
        ev = _PIN.get((str(device), 'h2d_event'))
        if ev is not None:
            ev.synchronize()        # the previous upload has left the staging buffer
        pin = _pinned(device, total, 'h2d')
        view = pin.numpy()
        off = 0
        for a in arrs:
            _host_copy(view[off:off + a.size], a.reshape(-1))
            off += a.size
        t = torch.empty(total, dtype=torch.float64, device=device)
        if total > _ZC_MAX or not _zero_copy_upload(pin, t, total):
            t.copy_(pin[:total], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        _PIN[(str(device), 'h2d_event')] = ev
    else:
        flat = np.concatenate([a.reshape(-1) for a in arrs]) if arrs else np.zeros(0)
        t = torch.from_numpy(flat).to(device)
    out, off = {}, 0
    for k, a in zip(keys, arrs):
        out[k] = t[off:off + a.size].reshape(a.shape if a.ndim > 0 else (1,))
        off += a.size
    return out


def to_host(flat):
    """One device->host copy of a flat fp64 tensor through a pinned buffer -> numpy (a view of the
    staging buffer: copy what you keep)."""
    if flat.is_cuda:
        n = flat.numel()
        pin = _pinned(flat.device, n, 'd2h')
        pin[:n].copy_(flat, non_blocking=True)
        torch.cuda.current_stream(flat.device).synchronize()
        return pin[:n].numpy()
    return flat.cpu().numpy()


def spd_inverse(A):
    """inverse and log-determinant of (a batch of) SPD matrices.
    Replaces np.linalg.inv / slogdet at base_models.py:464,471,476 and aep_models.py:68,78,91,525,533
    with the library's own kernel (one thread-block cluster per matrix, blocked Gauss-Jordan in
    fp64, `gpb_spd_inverse`): one launch, no host synchronisation; a non-SPD input yields NaNs,
    which the optimiser wrapper treats like the reference treats non-finite gradients."""
    return ops.spd_inverse(A.contiguous())


def bmv(A, x):
    """batched matrix-vector product [d,a,b] x [d,b] -> [d,a] (library kernel, tail.py)"""
    return tl.matvec(A, x)


def sandwich(L, X, alpha=1.0, C=None, beta=0.0):
    """alpha L X L (+ beta C) for (batches of) M x M matrices: two tensor-core products."""
    return tl.gemm(L, tl.gemm(X, L), alpha=alpha, C=C, beta=beta)


class Base_SGP_Layer(object):
    """base_models.py:168-658."""

    def __init__(self, no_train, input_size, output_size, no_pseudo, nat_param=True,
                 prec=None, device=None):
        self.Din, self.Dout, self.M, self.N = input_size, output_size, no_pseudo, no_train
        self.nat_param = nat_param
        self.prec = ops.PREC[prec if prec is not None else config.DEFAULT_PREC]
        self.device = device if device is not None else default_device()
        iu = torch.triu_indices(no_pseudo, no_pseudo)
        self._iu = (iu[0].to(self.device), iu[1].to(self.device))
        self._t = {}
        self._opnd = {}
        self._graphs = {}           # (fused alpha) -> TailGraph of the pre-tail (tailgraph.py)
        self._active_pre = None     # the captured pre-tail whose static outputs self._t holds
        self._cavity_done = None
        self.ls = np.zeros([input_size, ])
        self.sf = 0
        self.zu = np.zeros([no_pseudo, input_size])

    def __getattr__(self, name):
        if name in _EXPORTED:
            t = self.__dict__.get('_t', {})
            if name in self._B_SRC and self._B_SRC[name] in t:
                return self._B(name).detach().cpu().numpy()
            if name in t:
                return t[name].detach().cpu().numpy()
        raise AttributeError(name)

    # ---- hyper-parameter plumbing ---------------------------------------------------------
    def update_hypers(self, params, key_suffix='', _dev=None):
        """base_models.py:630-658: eta1_R -> R (log-diagonal upper triangle), theta_1 = R^T R.
        `_dev` (optional): the same dict already on the device (models upload all keys at once)."""
        dev = self.device
        self.ls = params['ls' + key_suffix]
        self.sf = params['sf' + key_suffix]
        self.zu = params['zu' + key_suffix]
        if _dev is None:
            _dev = pack_to_device({k: params[k + key_suffix] for k in ('ls', 'sf', 'zu', 'eta1_R', 'eta2')}, dev)
            key_suffix = ''
        ins = {k: _dev[k + key_suffix] for k in ('ls', 'sf', 'zu', 'eta1_R', 'eta2')}
        self._pre_tail(ins, getattr(self, '_fuse_cavity_alpha', None))

    @nvtx.annotate('pre_tail')
    def _pre_tail(self, ins, alpha):
        """The data-independent device work of one parameter update (+ cavity and log-partitions
        when the AEP objective announced its alpha): eager for the first calls, then one CUDA-graph
        replay per call (tailgraph.py)."""
        if len(self._graphs) > 4 and alpha not in self._graphs:
            self._graphs.clear()        # alpha is swept: do not hoard graph memory
        tg = self._graphs.setdefault(alpha, TailGraph())

        def fn(i):
            self._t = {}
            self._pre_device(i, alpha)
            return dict(self._t), self._cavity_done

        snap, done = tg.run(fn, ins, self.device)
        self._t = dict(snap)            # also drops the lazily formed B matrices of the last call
        self._opnd = {}
        self._cavity_done = done
        self._cavity_ready = None
        self._active_pre = tg if tg.captured else None

    def _pre_device(self, ins, alpha):
        M, Dout, dev = self.M, self.Dout, self.device
        t = self._t
        t['ls'] = ins['ls'].reshape(self.Din).contiguous()
        t['sf'] = ins['sf'].reshape(-1)[:1].contiguous()
        t['zu'] = ins['zu'].reshape(M, self.Din).contiguous()
        R = tl.unpack_r(ins['eta1_R'].reshape(Dout, -1).contiguous(), M)      # base_models.py:645-653
        t['theta_1_R'] = R
        t['theta_1'] = tl.gemm(R, R, ta=True)
        t['theta_2'] = ins['eta2'].reshape(Dout, M).contiguous()
        self._cavity_done = None
        self.compute_kuu()
        self.update_posterior()
        self._pre_extra(alpha)

    def _pre_extra(self, alpha):
        pass

    @nvtx.annotate('post_tail')
    def _post_tail(self, name, impl, st, *args):
        """Chain rules from the reduced statistics to the parameter gradients; replayed from a
        graph when this layer's state is the static output of a captured pre-tail."""
        pre = self._active_pre
        if pre is None:
            return impl(st, *args)
        tg = pre.children.setdefault((name,) + args, TailGraph(warmup=0))
        return tg.run(lambda i: impl(i, *args), {k: st[k] for k in _TAIL_STAT_KEYS}, self.device)

    def compute_kuu(self):
        """base_models.py:454-464."""
        t = self._t
        t['Kuu'] = ops.kmat(t['zu'], t['zu'], t['ls'], t['sf'], config.JITTER)
        t['Kuuinv'], t['logdet_Kuu'] = spd_inverse(t['Kuu'])

    def update_posterior(self):
        """base_models.py:466-488."""
        t = self._t
        Ki = t['Kuuinv']
        self._cavity_ready = None
        alpha = getattr(self, '_fuse_cavity_alpha', None)
        # log-determinants are kept as the kernel returns them: `ld_Su` holds log|Su^-1| when the inverse
        # was factorised (natural parameters) and log|Su| otherwise; _sign_Su turns it into log|Su|
        self._sign_Su = -1.0 if self.nat_param else 1.0
        if self.nat_param and alpha is not None:
            # AEP fast path: q(u) and the cavity need inv(Ki + theta_1) and inv(Ki + beta theta_1);
            # factorise both in ONE batched call (halves the launch count of the tail)
            beta = (self.N - alpha) * 1.0 / self.N
            Do = self.Dout
            both = torch.empty((2 * Do, self.M, self.M), dtype=_F, device=Ki.device)
            tl.lincomb([(1.0, Ki), (1.0, t['theta_1'])], out=both[:Do])
            tl.lincomb([(1.0, Ki), (beta, t['theta_1'])], out=both[Do:])
            inv, ld = spd_inverse(both)
            t['Suinv'], t['Suhatinv'] = both[:Do], both[Do:]
            t['Su'], t['Suhat'] = inv[:Do], inv[Do:]
            t['ld_Su'], t['ld_Suhat'] = ld[:Do], ld[Do:]
            t['mu'] = bmv(t['Su'], t['theta_2'])
            self._cavity_ready = alpha
        elif self.nat_param:
            t['Suinv'] = tl.lincomb([(1.0, Ki), (1.0, t['theta_1'])])
            t['Su'], t['ld_Su'] = spd_inverse(t['Suinv'])
            t['mu'] = bmv(t['Su'], t['theta_2'])
        else:
            t['Su'] = t['theta_1']
            t['Suinv'], t['ld_Su'] = spd_inverse(t['Su'])
            t['mu'] = t['theta_2']
        t['Splusmm'] = tl.lincomb([(1.0, t['Su'])], outer=(1.0, t['mu'], t['mu']))
        t['A'] = tl.matvec(Ki, t['mu'], t0=True)
        # B_sto = Ki Splusmm Ki - Ki and B_det = Ki Su Ki - Ki are formed on first use (_B): a layer
        # is fed either deterministic or uncertain inputs, never both in one objective call
        t.pop('B_sto', None)
        t.pop('B_det', None)
        self._opnd.pop('post', None)

    def get_hypers(self, key_suffix=''):
        """base_models.py:599-628."""
        M = self.M
        R = self.theta_1_R.copy()
        iu = np.triu_indices(M)
        di = np.diag_indices(M)
        eta1 = np.zeros((self.Dout, M * (M + 1) // 2))
        for d in range(self.Dout):
            Rd = R[d]
            Rd[di] = np.log(Rd[di])
            eta1[d, :] = Rd[iu]
        return {'ls' + key_suffix: self.ls, 'sf' + key_suffix: self.sf, 'zu' + key_suffix: self.zu,
                'eta1_R' + key_suffix: eta1, 'eta2' + key_suffix: self.theta_2}

    def init_hypers(self, x_train=None, key_suffix=''):
        """base_models.py:518-597 (host side: kmeans / median heuristic / random q(u))."""
        from scipy.cluster.vq import kmeans2
        from scipy.spatial.distance import cdist
        N, M, Din, Dout = self.N, self.M, self.Din, self.Dout
        if x_train is None:
            ls = np.log(np.ones((Din, )) + 0.1 * np.random.rand(Din, ))
            sf = np.log(np.array([1]))
            zu = np.tile(np.linspace(-1, 1, M).reshape((M, 1)), (1, Din))
        else:
            if N < 10000:
                centroids, label = kmeans2(x_train, M, minit='points')
            else:
                randind = np.random.permutation(N)
                centroids = x_train[randind[0:M], :]
            zu = centroids
            if N < 1000:
                X1 = np.copy(x_train)
            else:
                randind = np.random.permutation(N)
                X1 = x_train[randind[:1000], :]
            x_dist = cdist(X1, X1, 'euclidean')
            triu_ind = np.triu_indices(X1.shape[0])
            d2imed = np.median(x_dist[triu_ind])
            ls = np.log(d2imed / 2 + 1e-16) * np.ones((Din, ))
            sf = np.log(np.array([0.5]))
        ls2, sf2 = np.exp(2 * ls), np.exp(2 * sf)
        diff = zu[:, None, :] - zu[None, :, :]
        Kuu = sf2 * np.exp(-0.5 * np.sum(diff * diff / ls2, axis=2)) + config.JITTER * np.eye(M)
        Kuuinv = np.linalg.inv(Kuu)
        eta1_R = np.zeros((Dout, M * (M + 1) // 2))
        eta2 = np.zeros((Dout, M))
        iu = np.triu_indices(M)
        di = np.diag_indices(M)
        for d in range(Dout):
            mu = np.linspace(-1, 1, M).reshape((M, 1))
            alpha = 0.5 * np.random.rand(M)
            if self.nat_param:
                theta1 = np.diag(1 / alpha)
                theta2 = np.dot(theta1, mu)
            else:
                Su = np.linalg.inv(np.diag(1 / alpha) + Kuuinv)
                theta1 = Su
                theta2 = np.dot(Su, mu / alpha.reshape((M, 1)))
            R = np.linalg.cholesky(theta1).T
            R[di] = np.log(R[di])
            eta1_R[d, :] = R[iu]
            eta2[d, :] = theta2.reshape((M,))
        return {'sf' + key_suffix: sf, 'ls' + key_suffix: ls, 'zu' + key_suffix: zu,
                'eta1_R' + key_suffix: eta1_R, 'eta2' + key_suffix: eta2}

    # ---- device fast path ------------------------------------------------------------------
    def _AB(self, cav, stochastic):
        t = self._t
        if cav:
            return t['Ahat'], self._B('Bhat_sto' if stochastic else 'Bhat_det')
        return t['A'], self._B('B_sto' if stochastic else 'B_det')

    _B_SRC = {'B_sto': 'Splusmm', 'B_det': 'Su', 'Bhat_sto': 'Splusmmhat', 'Bhat_det': 'Suhat'}

    def _B(self, name):
        """base_models.py:485-488 / aep_models.py:541-546, lazily: Ki S Ki - Ki."""
        t = self._t
        if name not in t:
            Ki = t['Kuuinv']
            t[name] = sandwich(Ki, t[self._B_SRC[name]], C=Ki, beta=-1.0)
        return t[name]

    def _det_operands(self, cav):
        key = 'cav' if cav else 'post'
        if key not in self._opnd:
            A, B = self._AB(cav, False)
            self._opnd[key] = ops.DetOperands(self.prec, A.contiguous(), B.contiguous())
        return self._opnd[key]

    @nvtx.annotate('fwd_det')
    def _fwd_det(self, x, cav, save):
        """a5 on the device: (mout, vout, ctx)."""
        t = self._t
        opnd = self._det_operands(cav)
        m, v, Ks, Ts = ops.det_fwd(self.prec, x, t['zu'], t['ls'], t['sf'], opnd, save=save)
        return m, v, (x, opnd, Ks, Ts)

    @nvtx.annotate('bwd_det')
    def _bwd_det(self, ctx, dm, dv):
        """a8 per-row part: sufficient statistics of one deterministic layer."""
        t = self._t
        x, opnd, Ks, Ts = ctx
        dA, dzu, dl, dsf2 = ops.det_bwd(self.prec, x, t['zu'], t['ls'], t['sf'], opnd, dm, dv, Ks, Ts)
        dB = ops.det_syrk(self.prec, Ks, dv, self.M)
        return {'dA': dA, 'dB': dB, 'dzu': dzu, 'dl': dl, 'dsf2': dsf2, 'dvsum': tl.total(dv)}

    def det_chunks(self, n):
        """[(c0, c1)] row chunks of a deterministic-layer step whose saved Kfu / T buffers stay below
        config.DET_SAVE_BYTES (one chunk when everything fits)."""
        esize = 8 if self.prec == ops.F64 else 4
        MP = _lib.get().gpb_det_pad_m(self.M)
        per_row = (1 + self.Dout) * max(MP, 1) * esize
        rows = max(int(config.DET_SAVE_BYTES // per_row), config.DET_MIN_CHUNK_ROWS)
        if n <= rows:
            return [(0, n)]
        k = -(-n // rows)
        step = -(-n // k)
        if step >= 128:
            step = -(-step // 128) * 128      # whole row tiles
        return [(c0, min(c0 + step, n)) for c0 in range(0, n, step)]

    def det_step(self, xb, lik_fn, cav):
        """Forward, likelihood and per-row backward of a single deterministic layer, chunked over rows
        (aep_models.py:142-158 + lik + 452-493).  lik_fn(m, v, c0, c1) -> (dm, dv, extras: dict of additive
        [1]-tensors).  -> (statistics, summed extras)."""
        acc, ext = None, None
        for c0, c1 in self.det_chunks(xb.shape[0]):
            m, v, ctx = self._fwd_det(xb[c0:c1], cav=cav, save=True)
            dm, dv, e = lik_fn(m, v, c0, c1)
            st = self._bwd_det(ctx, dm, dv)
            del ctx, m, v
            if acc is None:
                acc, ext = st, e
            else:
                for d_acc, d_new in ((acc, st), (ext, e)):
                    for k in d_acc:
                        a = d_acc[k].reshape(1, -1)
                        tl.lincomb([(1.0, a), (1.0, d_new[k].reshape(1, -1))], out=a)
        return acc, ext

    @nvtx.annotate('fwd_mm')
    def _fwd_mm(self, mx, vx, cav, save=True):
        """a6 on the device (save=False: prediction, nothing kept for a backward)."""
        t = self._t
        A, B = self._AB(cav, True)
        m, v, vacc, psi1 = ops.mm_fwd(self.prec, mx, vx, t['zu'], t['ls'], t['sf'], A.contiguous(),
                                      B.contiguous(), save=save)
        return m, v, (mx, vx, cav, m, vacc, psi1)

    @nvtx.annotate('bwd_mm')
    def _bwd_mm(self, ctx, dm, dv):
        """a9 per-row part: statistics + per-row input gradients."""
        t = self._t
        mx, vx, cav, mout, vacc, psi1 = ctx
        A, B = self._AB(cav, True)
        return ops.mm_bwd(self.prec, mx, vx, t['zu'], t['ls'], t['sf'], A.contiguous(), B.contiguous(),
                          dm, dv, mout, vacc, psi1)

    @nvtx.annotate('fwd_mc')
    def _fwd_mc(self, mx, vx, eps, cav):
        """aep_models.py:160-180 / base_models.py:309-332: samples x = mx + sqrt(vx) eps (eps[K,n,Q]
        drawn on the host from numpy's global RNG, as the reference does) pushed through the
        deterministic-input kernels as K*n stacked rows.  -> mout, vout [K,n,Do], ctx."""
        K, n, Q = eps.shape
        xs = (eps * torch.sqrt(vx) + mx).reshape(K * n, Q).contiguous()
        m, v, ctx = self._fwd_det(xs, cav=cav, save=True)
        return m.reshape(K, n, self.Dout), v.reshape(K, n, self.Dout), (ctx, eps, vx)

    @nvtx.annotate('bwd_mc')
    def _bwd_mc(self, ctx, dm, dv):
        """Per-row part of backprop_grads_lvm_mc (aep_models.py:307-410, vfe_models.py:405-476) +
        backprop_grads_reparam (base_models.py:373-388): the deterministic-layer statistics over the
        stacked samples and the gradient wrt every sample, folded back onto (mx, vx)."""
        det_ctx, eps, vx = ctx
        K, n, Q = eps.shape
        dm2 = dm.reshape(K * n, self.Dout).contiguous()
        dv2 = dv.reshape(K * n, self.Dout).contiguous()
        st = self._bwd_det(det_ctx, dm2, dv2)
        xs, opnd, Ks, Ts = det_ctx
        t = self._t
        dx = ops.det_dx(self.prec, xs, t['zu'], t['ls'], opnd, dm2, dv2, Ks, Ts).reshape(K, n, Q)
        st['dmx'] = dx.sum(0)
        st['dvx'] = (dx * eps).sum(0) / (2.0 * torch.sqrt(vx))
        return st

    def _forward_mc_iface(self, mx, vx, cav):
        """Layer-level Monte-Carlo forward (aep_models.py:160-180 / base_models.py:309-332): eps from
        numpy's global RNG as in the reference, samples through the deterministic-input kernels.
        -> (mout, vout, kfu, x, eps) [K,n,.] and their stacked [K*n,.] views, numpy."""
        dev, t = self.device, self._t
        n = mx.shape[0]
        eps_h = np.random.randn(config.MC_NO_SAMPLES, n, self.Din)
        m, v, (det_ctx, eps, _) = self._fwd_mc(to_dev(mx, dev), to_dev(vx, dev), to_dev(eps_h, dev), cav=cav)
        xs = det_ctx[0]
        K = config.MC_NO_SAMPLES
        m_stk = m.reshape(K * n, self.Dout).cpu().numpy()
        v_stk = v.reshape(K * n, self.Dout).cpu().numpy()
        kfu_stk = ops.kmat(xs, t['zu'], t['ls'], t['sf']).cpu().numpy()
        x_stk = xs.cpu().numpy()
        e_stk = eps_h.reshape(K * n, self.Din)
        return ((m_stk.reshape(K, n, self.Dout), v_stk.reshape(K, n, self.Dout), kfu_stk.reshape(K, n, self.M),
                 x_stk.reshape(K, n, self.Din), eps_h), (m_stk, v_stk, kfu_stk, x_stk, e_stk))

    def _backprop_mc_iface(self, dm, dv, x, cav):
        """Statistics + per-sample input gradient of the stacked samples (Kfu regenerated on chip)."""
        dev, t = self.device, self._t
        xs = to_dev(x, dev)
        n = xs.shape[0]
        _, _, ctx = self._fwd_det(xs, cav=cav, save=True)
        dm2 = to_dev(np.reshape(dm, (n, self.Dout)), dev)
        dv2 = to_dev(np.reshape(dv, (n, self.Dout)), dev)
        st = self._bwd_det(ctx, dm2, dv2)
        _, opnd, Ks, Ts = ctx
        return st, ops.det_dx(self.prec, xs, t['zu'], t['ls'], opnd, dm2, dv2, Ks, Ts)

    def backprop_grads_reparam(self, dx, m, v, eps):
        """base_models.py:373-388: sample gradients folded back onto the input mean / variance."""
        dev = self.device
        e = to_dev(eps, dev)
        d = to_dev(dx, dev).reshape(e.shape)
        return {'mx': d.sum(0).cpu().numpy(),
                'vx': ((d * e).sum(0) / (2.0 * torch.sqrt(to_dev(v, dev)))).cpu().numpy()}

    def _predictive_dx(self, x, dm_dm, dm_dv):
        """d(sum_d dm_dm*m_d + dm_dv*v_d)/dx of the posterior deterministic-input layer on the device:
        det_fwd (saving Kfu and T = B_det kfu) followed by the det_dx kernel.  -> dx[n,D]."""
        t = self._t
        m, v, ctx = self._fwd_det(x, cav=False, save=True)
        _, opnd, Ks, Ts = ctx
        n = x.shape[0]
        dm = to_dev(np.asarray(dm_dm, dtype=np.float64), self.device).expand(n, self.Dout).contiguous()
        dv = to_dev(np.asarray(dm_dv, dtype=np.float64), self.device).expand(n, self.Dout).contiguous()
        return m, v, ops.det_dx(self.prec, x, t['zu'], t['ls'], opnd, dm, dv, Ks, Ts)

    def backprop_predictive_grads_reg(self, m, v, dm_dm, dm_dv, dv_dm, dv_dv, kfu, x):
        """base_models.py:391-426.  The reference forms the second gradient from ``dkfu_m`` as well
        (base_models.py:424-425), so both returned arrays are d m / d x; kept as is (dv_dm / dv_dv
        are accepted and, as there, unused).  Kfu is regenerated on the device from ``x``."""
        _, _, dx = self._predictive_dx(to_dev(x, self.device), dm_dm, dm_dv)
        dx = dx.cpu().numpy()
        return dx, dx.copy()

    # ---- shared chain rules ------------------------------------------------------------------
    def _pack_eta1(self, terms):
        """theta_1 = R^T R with log-diagonal packing (base_models.py:505-514) of sum_i c_i dtheta1_i;
        terms: one or two (coef, [Dout,M,M])."""
        R = self._t['theta_1_R']
        sym = tl.lincomb([(c, X) for c, X in terms] + [(c, X, True) for c, X in terms])
        return tl.pack_r(tl.gemm(R, sym), R)

    def _posterior_grad_u(self, dmu, dSu):
        """base_models.py:490-516 -> (dSuinv or dSu [the theta_1 gradient before packing], deta2, dKuuinv or None)."""
        t = self._t
        if self.nat_param:
            dSu = tl.lincomb([(1.0, dSu)], outer=(1.0, dmu, t['theta_2']))
            dSuinv = sandwich(t['Su'], dSu, alpha=-1.0)
            return dSuinv, bmv(t['Su'], dmu), tl.lincomb([(1.0, dSuinv)], reduce=True)
        return dSu, dmu, None

    def compute_posterior_grad_u(self, dmu, dSu):
        d1, e2, dKi = self._posterior_grad_u(to_dev(dmu, self.device), to_dev(dSu, self.device))
        e1 = self._pack_eta1([(1.0, d1)])
        dKi = dKi.cpu().numpy() if dKi is not None else np.zeros((self.M, self.M))
        return e1.cpu().numpy(), e2.cpu().numpy(), dKi

    def _stats_record(self, st):
        """[dzu | dl | dsf2 | dvsum] as one contiguous record (the order of a layer's packed statistics)."""
        parts = [st['dzu'], st['dl'], st['dsf2'], st['dvsum']]
        p0 = parts[0].data_ptr()
        off = 0
        for x in parts:
            if x.data_ptr() != p0 + 8 * off or not x.is_contiguous():
                return tl.gather([q.contiguous() for q in parts])
            off += x.numel()
        return torch.as_strided(parts[0], (off,), (1,))

    def _kernel_hyper_tail(self, st, Mm):
        """aep_models.py:455-460 + 497-504 + kernels.py:447-475 (d_trace_MKzz_dhypers with
        Kzz = Kuu - JITTER I): fold direct kernel derivatives with the Kuu path; one launch."""
        t = self._t
        M, D = self.M, self.Din
        out = tl.khyper(Mm, t['Kuu'], t['zu'], t['ls'], t['sf'], self._stats_record(st), config.JITTER)
        return out[0:1], out[1:1 + D], out[1 + D:].reshape(M, D)

    def sample(self, x):
        """base_models.py:428-452: one joint draw of f(x) -- u ~ q(u), then f | u at the test
        inputs.  The standard-normal draws come from numpy's global RNG in the reference's order;
        the algebra (two Cholesky factorisations, kernel matrices from the library) runs on the
        device."""
        t = self._t
        dev = self.device
        xd = to_dev(x, dev)
        Lu = torch.linalg.cholesky(t['Su'])
        eps_u = to_dev(np.random.randn(self.Dout, self.M), dev)
        u = t['mu'] + torch.matmul(Lu, eps_u.unsqueeze(-1)).squeeze(-1)
        n = xd.shape[0]
        kff = ops.kmat(xd, xd, t['ls'], t['sf'], config.JITTER)
        kfu = ops.kmat(xd, t['zu'], t['ls'], t['sf'])
        qfu = torch.matmul(kfu, t['Kuuinv'])
        mf = torch.matmul(qfu, u.t())
        vf = kff - torch.matmul(qfu, kfu.t())
        Lf = torch.linalg.cholesky(vf)
        eps_f = to_dev(np.random.randn(n, self.Dout), dev)
        return (mf + torch.matmul(Lf, eps_f)).cpu().numpy()

    # ---- prediction path (base_models.py:235-307) -----------------------------------------
    def forward_prop_thru_post(self, mx, vx=None, mode=config.PROP_MM, return_info=False):
        dev = self.device
        if vx is None:
            x = to_dev(mx, dev)
            m, v, _ = self._fwd_det(x, cav=False, save=False)
            if return_info:
                t = self._t
                return m.cpu().numpy(), v.cpu().numpy(), ops.kmat(x, t['zu'], t['ls'], t['sf']).cpu().numpy()
            return m.cpu().numpy(), v.cpu().numpy()
        if mode == config.PROP_MM:
            a, b = to_dev(mx, dev), to_dev(vx, dev)
            m, v, _ = self._fwd_mm(a, b, cav=False, save=False)
            if return_info:
                t = self._t
                p1, p2 = ops.psi_stats(a, b, t['zu'], t['ls'], t['sf'])
                return m.cpu().numpy(), v.cpu().numpy(), p1.cpu().numpy(), p2.cpu().numpy()
            return m.cpu().numpy(), v.cpu().numpy()
        if mode == config.PROP_MC:
            res, res_s = self._forward_mc_iface(mx, vx, cav=False)
            return (res, res_s) if return_info else (res[0], res[1])
        if mode == config.PROP_LIN:
            raise NotImplementedError('Prediction with linearisation not implemented TODO')   # base_models.py:255-259
        raise NotImplementedError('unknown propagation mode')


class AEP_SGP_Layer(Base_SGP_Layer):
    """aep_models.py:26-586."""

    def compute_cavity(self, alpha):
        """aep_models.py:513-546."""
        if self._cavity_done is not None and self._cavity_done == alpha:
            return                      # formed with the parameter update (_pre_extra)
        self._cavity_done = None
        t = self._t
        Ki = t['Kuuinv']
        beta = (self.N - alpha) * 1.0 / self.N
        if self.nat_param and getattr(self, '_cavity_ready', None) == alpha:
            t['muhat'] = tl.matvec(t['Suhat'], t['theta_2'], c0=beta)     # Suhat came with the posterior
        elif self.nat_param:
            t['Suhatinv'] = tl.lincomb([(1.0, Ki), (beta, t['theta_1'])])
            t['Suhat'], t['ld_Suhat'] = spd_inverse(t['Suhatinv'])
            t['muhat'] = tl.matvec(t['Suhat'], t['theta_2'], c0=beta)
        else:
            f2 = bmv(t['Suinv'], t['mu'])
            t['Suhatinv'] = tl.lincomb([(1.0 - beta, Ki), (beta, t['Suinv'])])      # Ki + beta (Suinv - Ki)
            t['Suhat'], t['ld_Suhat'] = spd_inverse(t['Suhatinv'])
            t['muhat'] = tl.matvec(t['Suhat'], f2, c0=beta)
        t['Ahat'] = tl.matvec(Ki, t['muhat'], t0=True)
        t['Splusmmhat'] = tl.lincomb([(1.0, t['Suhat'])], outer=(1.0, t['muhat'], t['muhat']))
        t.pop('Bhat_sto', None)      # formed on first use (_B)
        t.pop('Bhat_det', None)
        t.pop('phi', None)
        self._opnd.pop('cav', None)

    def _pre_extra(self, alpha):
        """AEP objective: cavity and log-partitions belong to the same (graphable) phase."""
        if alpha is not None:
            self.compute_cavity(alpha)
            self._t['phi'] = self._phi(alpha)
            self._cavity_done = alpha

    def _phi(self, alpha):
        """aep_models.py:62-114 (device scalar)."""
        t = self._t
        if 'phi' in t and self._cavity_done is not None and self._cavity_done == alpha:
            return t['phi']
        N = self.N
        sp, sc = N * 1.0 / alpha - 1.0, N * 1.0 / alpha
        v1 = bmv(t['Suinv'], t['mu'])
        v2 = bmv(t['Suhatinv'], t['muhat'])
        # phi_prior + sp phi_post - sc phi_cav; log|Suhat| = -ld_Suhat (the inverse was factorised)
        return tl.dots([(0.5 * self.Dout, t['logdet_Kuu'], None), (0.5 * sp * self._sign_Su, t['ld_Su'], None),
                        (0.5 * sp, t['mu'], v1), (0.5 * sc, t['ld_Suhat'], None), (-0.5 * sc, t['muhat'], v2)])

    def compute_phi(self, alpha=1.0):
        return float(self._phi(alpha).item())

    def compute_phi_prior(self):
        """aep_models.py:62-70."""
        return float(0.5 * self.Dout * self._t['logdet_Kuu'].item())

    def compute_phi_posterior(self):
        """aep_models.py:72-83."""
        t = self._t
        return float(tl.dots([(0.5 * self._sign_Su, t['ld_Su'], None),
                              (0.5, t['mu'], bmv(t['Suinv'], t['mu']))]).item())

    def compute_phi_cavity(self):
        """aep_models.py:85-97 (after compute_cavity)."""
        t = self._t
        return float(tl.dots([(-0.5, t['ld_Suhat'], None),
                              (0.5, t['muhat'], bmv(t['Suhatinv'], t['muhat']))]).item())

    def _cav_grad_u(self, dmu, dSu, alpha):
        """aep_models.py:548-586 -> (theta_1 gradient before packing, deta2, dKuuinv)."""
        t = self._t
        beta = (self.N - alpha) * 1.0 / self.N
        if self.nat_param:
            dSu = tl.lincomb([(1.0, dSu)], outer=(beta, dmu, t['theta_2']))
            dSuinv = sandwich(t['Suhat'], dSu, alpha=-1.0)
            return ((beta, dSuinv),), tl.matvec(t['Suhat'], dmu, c0=beta), tl.lincomb([(1.0, dSuinv)], reduce=True)
        f2 = bmv(t['Suinv'], t['mu'])
        dSuhat = tl.lincomb([(1.0, dSu)], outer=(beta, dmu, f2))
        dSuhatinv = sandwich(t['Suhat'], dSuhat, alpha=-1.0)
        Sdm = bmv(t['Suhat'], dmu)
        dSuinv = tl.lincomb([(beta, dSuhatinv)], outer=(beta, Sdm, t['mu']))
        dtheta1 = sandwich(t['Suinv'], dSuinv, alpha=-1.0)
        return ((1.0, dtheta1),), tl.matvec(t['Suinv'], Sdm, c0=beta), \
            tl.lincomb([(1.0 - beta, dSuhatinv)], reduce=True)

    def compute_cav_grad_u(self, dmu, dSu, alpha):
        d1, e2, dKi = self._cav_grad_u(to_dev(dmu, self.device), to_dev(dSu, self.device), alpha)
        return self._pack_eta1(list(d1)).cpu().numpy(), e2.cpu().numpy(), dKi.cpu().numpy()

    def _tail_det(self, st, alpha):
        return self._post_tail('det', self._tail_det_impl, st, alpha)

    def _tail_mm(self, st, alpha):
        return self._post_tail('mm', self._tail_mm_impl, st, alpha)

    def _tail_mc(self, st, alpha):
        return self._post_tail('mc', self._tail_mc_impl, st, alpha)

    def _dKi_common(self, dA, dB, S):
        """sum_d dA_d muhat_d^T + 2 sum_d (Ki S_d)^T dB_d as tensor-core products with K = Dout (resp.
        Dout M): the sums over the output dimensions are the products' inner dimension."""
        t = self._t
        Do, M = self.Dout, self.M
        Y = tl.gemm(t['Kuuinv'], S)                                                # [Do, M, M]
        Z = tl.gemm(Y.reshape(Do * M, M), dB.reshape(Do * M, M), ta=True, alpha=2.0)
        return tl.gemm(dA, t['muhat'], ta=True, C=Z, beta=1.0)

    def _tail_mc_impl(self, st, alpha):
        """aep_models.py:352-403 (backprop_grads_lvm_mc) on the statistics of the stacked samples
        (dA = sum_n dm kfu, dB = sum_n dv kfu kfu^T): with SK = Suhat Kuuinv,
        dSinv_d = -SK dB_d SK^T - (SK dA_d) muhat_d^T, dtheta2_d = beta SK dA_d + phi terms."""
        t = self._t
        N, Ki = self.N, t['Kuuinv']
        Do, M = self.Dout, self.M
        if not self.nat_param:
            raise NotImplementedError('AEP Monte-Carlo propagation needs nat_param=True (the reference '
                                      'treats theta as natural parameters: aep_models.py:365-371)')
        beta = (N - alpha) * 1.0 / N
        scale_post = N * 1.0 / alpha - 1.0
        scale_cav = -N * 1.0 / alpha
        dA, dB = st['dA'], st['dB']
        SK = tl.gemm(t['Suhat'], Ki)
        SKdA = bmv(SK, dA)
        X = tl.gemm(SK, tl.gemm(dB, SK, tb=True))
        dSinv = tl.lincomb([(-1.0, X)], outer=(-1.0, SKdA, t['muhat']))
        dtheta1 = tl.lincomb([(beta, dSinv), (-0.5 * scale_post, t['Splusmm']),
                              (-0.5 * scale_cav * beta, t['Splusmmhat'])])
        dtheta2 = tl.veccomb([(beta, SKdA), (scale_post, t['mu']), (scale_cav * beta, t['muhat'])])
        # dKi = dA^T muhat + 2 sum_d SK_d dB_d - sum_d dB_d + sum_d dSinv_d
        Z = tl.gemm(self._stack_t(SK), dB.reshape(Do * M, M), ta=True, alpha=2.0)
        Z = tl.gemm(dA, t['muhat'], ta=True, C=Z, beta=1.0)
        T1 = tl.lincomb([(-1.0, dB), (1.0, dSinv), (-0.5 * scale_post, t['Splusmm']),
                         (-0.5 * scale_cav, t['Splusmmhat'])], reduce=True)
        # Minner = scale_post sum Splusmm + scale_cav sum Splusmmhat - 2 dKi  = -2 (Z + T1)
        Minner = tl.lincomb([(-2.0, Z), (-2.0, T1)])
        M_all = sandwich(Ki, Minner, alpha=0.5, C=Ki, beta=0.5 * Do)
        dsf, dls, dzu = self._kernel_hyper_tail(st, M_all)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': self._pack_eta1([(1.0, dtheta1)]), 'eta2': dtheta2}

    def _stack_t(self, X):
        """[Do, M, M] -> the [Do M, M] matrix of the TRANSPOSED blocks (so that a product with ta=True
        contracts sum_d X_d Y_d)."""
        return tl.lincomb([(1.0, X, True)]).reshape(self.Dout * self.M, self.M)

    def _tail_det_impl(self, st, alpha):
        """aep_models.py:462-511 rewritten on the statistics (dA = sum_n dm kfu,
        dB = sum_n dv kfu kfu^T): dmucav = dA Kuuinv, dSucav = Kuuinv dB Kuuinv."""
        t = self._t
        N, Ki = self.N, t['Kuuinv']
        Do = self.Dout
        scale_post = N * 1.0 / alpha - 1.0
        scale_cav = -N * 1.0 / alpha
        dA, dB = st['dA'], st['dB']
        Sim = bmv(t['Suhatinv'], t['muhat'])
        dmucav = tl.matvec(Ki, dA, t0=True, w0=Sim, cw0=scale_cav)
        dSucav = tl.lincomb([(1.0, sandwich(Ki, dB)), (0.5 * scale_cav, t['Suhatinv'])],
                            outer=(-0.5 * scale_cav, Sim, Sim))
        d1c, e2c, dKi_cav = self._cav_grad_u(dmucav, dSucav, alpha)
        Simp = bmv(t['Suinv'], t['mu'])
        dSu = tl.lincomb([(0.5 * scale_post, t['Suinv'])], outer=(-0.5 * scale_post, Simp, Simp))
        d1p, e2p, dKi_post = self._posterior_grad_u(tl.veccomb([(scale_post, Simp)]), dSu)
        # dKi = dA^T muhat + 2 sum_d (Ki Suhat_d)^T dB_d - sum_d dB_d + dKi_cav + dKi_post - Dout/2 Kuu
        Z = self._dKi_common(dA, dB, t['Suhat'])
        sumB = tl.lincomb([(1.0, dB)], reduce=True)
        terms = [(1.0, Z), (-1.0, sumB), (1.0, dKi_cav)] + ([(1.0, dKi_post)] if dKi_post is not None else [])
        dKi = tl.lincomb(terms)
        dKi = tl.lincomb([(1.0, dKi), (-0.5 * Do, t['Kuu'])])
        Mm = sandwich(Ki, dKi, alpha=-1.0)
        dsf, dls, dzu = self._kernel_hyper_tail(st, Mm)
        e1 = self._pack_eta1(list(d1c) + [(1.0, d1p)])
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': e1, 'eta2': tl.veccomb([(1.0, e2c), (1.0, e2p)])}

    def _tail_mm_impl(self, st, alpha):
        """aep_models.py:252-297 on the statistics (dA = sum_n dm_all psi1, dB = sum_n dv psi2)."""
        t = self._t
        N, Ki = self.N, t['Kuuinv']
        Do = self.Dout
        if not self.nat_param:
            raise NotImplementedError('AEP moment-matched layers need nat_param=True (the reference '
                                      'has no valid non-natural variant: aep_models.py:252-297)')
        beta = (N - alpha) * 1.0 / N
        scale_post = N * 1.0 / alpha - 1.0
        scale_cav = -N * 1.0 / alpha
        dA, dB = st['dA'], st['dB']
        dvcav = sandwich(Ki, dB)
        dmcav = tl.matvec(dvcav, t['muhat'], c0=2.0, A1=Ki, x1=dA, t1=True)
        dvcav = tl.lincomb([(1.0, dvcav)], outer=(beta, dmcav, t['theta_2']))
        dvcavinv = sandwich(t['Suhat'], dvcav, alpha=-1.0)
        dtheta1 = tl.lincomb([(beta, dvcavinv), (-0.5 * scale_post, t['Splusmm']),
                              (-0.5 * scale_cav * beta, t['Splusmmhat'])])
        dtheta2 = tl.matvec(t['Suhat'], dmcav, c0=beta, w0=t['mu'], cw0=scale_post, w1=t['muhat'],
                            cw1=scale_cav * beta)
        # dKi = dA^T muhat + 2 sum_d (Ki Splusmmhat_d)^T dB_d - sum_d dB_d + sum_d dvcavinv_d
        Z = self._dKi_common(dA, dB, t['Splusmmhat'])
        # Minner = scale_post sum Splusmm + scale_cav sum Splusmmhat - 2 dKi
        T1 = tl.lincomb([(2.0, dB), (-2.0, dvcavinv), (scale_post, t['Splusmm']), (scale_cav, t['Splusmmhat'])],
                        reduce=True)
        Minner = tl.lincomb([(-2.0, Z), (1.0, T1)])
        M_all = sandwich(Ki, Minner, alpha=0.5, C=Ki, beta=0.5 * Do)
        dsf, dls, dzu = self._kernel_hyper_tail(st, M_all)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': self._pack_eta1([(1.0, dtheta1)]), 'eta2': dtheta2}

    # ---- layer-level API of the reference (numpy in / numpy out; small n) ------------------
    def forward_prop_thru_cav(self, mx, vx=None, mode=config.PROP_MM):
        """aep_models.py:116-140; returns the materialised kfu / psi like the reference."""
        dev, t = self.device, self._t
        if vx is None:
            x = to_dev(mx, dev)
            m, v, _ = self._fwd_det(x, cav=True, save=False)
            return m.cpu().numpy(), v.cpu().numpy(), ops.kmat(x, t['zu'], t['ls'], t['sf']).cpu().numpy()
        if mode == config.PROP_MM:
            a, b = to_dev(mx, dev), to_dev(vx, dev)
            m, v, _ = self._fwd_mm(a, b, cav=True, save=False)
            p1, p2 = ops.psi_stats(a, b, t['zu'], t['ls'], t['sf'])
            return m.cpu().numpy(), v.cpu().numpy(), p1.cpu().numpy(), p2.cpu().numpy()
        if mode == config.PROP_MC:
            return self._forward_mc_iface(mx, vx, cav=True)
        if mode == config.PROP_LIN:
            raise NotImplementedError('prop_mode LIN: the reference has no linearised layer either '
                                      '(aep_models.py:135-136 calls an undefined method)')
        raise NotImplementedError('unknown propagation mode')

    def backprop_grads_lvm_mc(self, m, v, dm, dv, kfu, x, alpha=1.0):
        """aep_models.py:307-410 on stacked samples x[K*n,Din]: -> (hyper grads, dx[K*n,Din])."""
        st, dx = self._backprop_mc_iface(dm, dv, x, cav=True)
        return {k: g.cpu().numpy() for k, g in self._tail_mc(st, alpha).items()}, dx.cpu().numpy()

    def backprop_grads_reg(self, m, v, dm, dv, kfu, x, alpha=1.0):
        """aep_models.py:413-511.  kfu is recomputed on chip; the argument is ignored."""
        dev = self.device
        xd = to_dev(x, dev)
        _, _, ctx = self._fwd_det(xd, cav=True, save=True)
        st = self._bwd_det(ctx, to_dev(dm, dev), to_dev(dv, dev))
        return {k: g.cpu().numpy() for k, g in self._tail_det(st, alpha).items()}

    def backprop_grads_lvm_mm(self, m, v, dm, dv, psi1, psi2, mx, vx, alpha=1.0):
        """aep_models.py:202-304.  psi1 / psi2 (and m, v: the caller's v may already carry the
        likelihood's in-place noise term, lik_layers.py:121) are regenerated on chip."""
        dev = self.device
        _, _, ctx = self._fwd_mm(to_dev(mx, dev), to_dev(vx, dev), cav=True)
        st = self._bwd_mm(ctx, to_dev(dm, dev), to_dev(dv, dev))
        gh = {k: g.cpu().numpy() for k, g in self._tail_mm(st, alpha).items()}
        return gh, {'mx': st['dmx'].cpu().numpy(), 'vx': st['dvx'].cpu().numpy()}


class VFE_SGP_Layer(Base_SGP_Layer):
    """vfe_models.py:290-548."""

    def _pre_extra(self, alpha):
        self._t['kl'] = self._kl()

    def _kl(self):
        """vfe_models.py:309-325."""
        t = self._t
        if 'kl' in t:
            return t['kl']
        Ssum = tl.lincomb([(1.0, t['Splusmm'])], reduce=True)
        # 0.5 (Dout log|Kuu| - sum log|Su| - Dout M + tr(Kuuinv Splusmm))
        return tl.dots([(0.5 * self.Dout, t['logdet_Kuu'], None), (-0.5 * self._sign_Su, t['ld_Su'], None),
                        (0.5, t['Kuuinv'], Ssum)], const=-0.5 * self.Dout * self.M)

    def compute_KL(self):
        return float(self._kl().item())

    def _tail(self, st, stochastic):
        return self._post_tail('vfe', self._tail_impl, st, stochastic)

    def _tail_impl(self, st, stochastic):
        """vfe_models.py:518-541 (det) / 363-394 (mm) on the statistics."""
        t = self._t
        Ki = t['Kuuinv']
        Do, M = self.Dout, self.M
        dA, dB = st['dA'], st['dB']
        dSu0 = sandwich(Ki, dB)
        # dmu = dA Ki (+ 2 dSu mu) + mu Ki
        if stochastic:
            dmu = tl.matvec(Ki, dA, t0=True, A1=dSu0, x1=t['mu'], c1=2.0, w0=t['A'])
        else:
            dmu = tl.matvec(Ki, dA, t0=True, w0=t['A'])
        dSu = tl.lincomb([(1.0, dSu0), (0.5, Ki), (-0.5, t['Suinv'])])
        d1, e2, dKi_u = self._posterior_grad_u(dmu, dSu)
        S = t['Splusmm'] if stochastic else t['Su']
        Y = tl.gemm(Ki, S)
        Z = tl.gemm(Y.reshape(Do * M, M), dB.reshape(Do * M, M), ta=True, alpha=2.0)
        Z = tl.gemm(dA, t['mu'], ta=True, C=Z, beta=1.0)
        T1 = tl.lincomb([(-1.0, dB), (0.5, t['Splusmm'])], reduce=True)
        terms = [(1.0, Z), (1.0, T1), (-0.5 * Do, t['Kuu'])] + ([(1.0, dKi_u)] if dKi_u is not None else [])
        dKi = tl.lincomb(terms)
        Mm = sandwich(Ki, dKi, alpha=-1.0)
        dsf, dls, dzu = self._kernel_hyper_tail(st, Mm)
        return {'sf': dsf, 'ls': dls, 'zu': dzu, 'eta1_R': self._pack_eta1([(1.0, d1)]), 'eta2': e2}

    def backprop_grads_reg(self, m, v, dm, dv, kfu, x):
        """vfe_models.py:479-548."""
        dev = self.device
        _, _, ctx = self._fwd_det(to_dev(x, dev), cav=False, save=True)
        st = self._bwd_det(ctx, to_dev(dm, dev), to_dev(dv, dev))
        return {k: g.cpu().numpy() for k, g in self._tail(st, False).items()}

    def backprop_grads_lvm_mc(self, m, v, dm, dv, kfu, x):
        """vfe_models.py:405-476 on stacked samples: -> (hyper grads, dx[K*n,Din])."""
        st, dx = self._backprop_mc_iface(dm, dv, x, cav=False)
        return {k: g.cpu().numpy() for k, g in self._tail(st, False).items()}, dx.cpu().numpy()

    def backprop_grads_lvm_mm(self, m, v, dm, dv, psi1, psi2, mx, vx):
        """vfe_models.py:328-401."""
        dev = self.device
        _, _, ctx = self._fwd_mm(to_dev(mx, dev), to_dev(vx, dev), cav=False)
        st = self._bwd_mm(ctx, to_dev(dm, dev), to_dev(dv, dev))
        gh = {k: g.cpu().numpy() for k, g in self._tail(st, True).items()}
        return gh, {'mx': st['dmx'].cpu().numpy(), 'vx': st['dvx'].cpu().numpy()}
