"""Shared bodies of the kernel-level parity checks.  tests/test_emu_ops.py runs them on the
CPU fiber emulator, tests/test_gpu_ops.py on the real CUDA library (DEV = 'cuda')."""
import numpy as np
import torch

import geepee_oracle as go
import golden_util as gu

DEV = 'cpu'


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def problem(n, M, D, Do, seed, uncertain=False):
    rng = np.random.RandomState(seed)
    p = dict(x=rng.standard_normal((n, D)), z=rng.standard_normal((M, D)),
             ls=0.3 * rng.standard_normal(D) + 0.3, sf=np.array([0.2]),
             A=rng.standard_normal((Do, M)), dm=rng.standard_normal((n, Do)),
             dv=rng.standard_normal((n, Do)))
    B = rng.standard_normal((Do, M, M))
    p['B'] = B + np.transpose(B, (0, 2, 1))
    if uncertain:
        p['vx'] = 0.05 + rng.rand(n, D)
        p['Bn'] = B                      # non-symmetric B exercises the symmetrisation
    return p


DET_SHAPES = [(37, 5, 2, 3), (300, 10, 1, 1), (150, 130, 3, 2), (70, 260, 17, 1), (33, 128, 9, 9)]


def check_det_layer(n, M, D, Do, prec, tol):
    from geepee_b200 import ops
    pr = ops.PREC[prec]
    p = problem(n, M, D, Do, seed=n + M)
    x, z, ls, sf = T(p['x']), T(p['z']), T(p['ls']), T(p['sf'])
    opnd = ops.DetOperands(pr, T(p['A']), T(p['B']))
    mout, vout, Ks, Ts = ops.det_fwd(pr, x, z, ls, sf, opnd, save=True)
    kfu = go.ard_kernel(2 * p['ls'], 2 * p['sf'], p['x'], p['z'])
    m_ref = np.einsum('nm,dm->nd', kfu, p['A'])
    v_ref = np.exp(2 * p['sf']) + np.einsum('dab,na,nb->nd', p['B'], kfu, kfu)
    assert gu.rel_err(N(mout), m_ref) < tol
    assert gu.rel_err(N(vout), v_ref) < tol
    assert gu.rel_err(N(Ks)[:, :M], kfu) < tol
    assert np.all(N(Ks)[:, M:] == 0)
    assert gu.rel_err(N(Ts)[:, :, :M], np.einsum('dab,nb->nda', p['B'], kfu)) < tol
    # backward
    dm, dv = T(p['dm']), T(p['dv'])
    dA, dzu, dl, dsf2 = ops.det_bwd(pr, x, z, ls, sf, opnd, dm, dv, Ks, Ts)
    dB = ops.det_syrk(pr, Ks, dv, M)
    dkfu = np.einsum('nd,dm->nm', p['dm'], p['A']) + 2 * np.einsum('nd,dab,na->nb', p['dv'], p['B'], kfu)
    r_var, r_l, r_z = go.kfu_derivs(dkfu, kfu, np.exp(p['ls']), np.exp(2 * p['sf']), p['x'], p['z'])
    assert gu.rel_err(N(dA), np.einsum('nd,nm->dm', p['dm'], kfu)) < tol
    assert gu.rel_err(N(dB), np.einsum('nd,na,nb->dab', p['dv'], kfu, kfu)) < tol
    assert gu.rel_err(N(dzu), r_z) < tol
    assert gu.rel_err(N(dl), r_l) < tol
    assert gu.rel_err(N(dsf2), np.ravel(r_var)) < tol
    # input gradient (Monte-Carlo propagation; kernels.py:393-395 kfucompDer grad_x=True)
    dx = ops.det_dx(pr, x, z, ls, opnd, dm, dv, Ks, Ts)
    r_x = go.kfu_derivs(dkfu, kfu, np.exp(p['ls']), np.exp(2 * p['sf']), p['x'], p['z'], grad_x=True)[3]
    assert gu.rel_err(N(dx), r_x) < tol


MM_SHAPES = [(9, 6, 3, 2), (40, 5, 2, 1), (21, 50, 1, 4), (13, 12, 5, 3), (11, 7, 7, 2), (10, 9, 4, 6),
             (12, 8, 3, 12), (9, 10, 2, 20), (8, 7, 5, 40), (35, 6, 16, 3), (14, 11, 4, 4), (16, 9, 3, 3),
             (150, 12, 5, 50), (70, 9, 8, 9), (45, 13, 1, 64), (33, 6, 16, 7)]


def check_mm_layer(n, M, Q, Do, prec, tol):
    from geepee_b200 import ops
    pr = ops.PREC[prec]
    p = problem(n, M, Q, Do, seed=7 * n + M, uncertain=True)
    mx, vx, z, ls, sf = T(p['x']), T(p['vx']), T(p['z']), T(p['ls']), T(p['sf'])
    A, B = T(p['A']), T(p['Bn'])
    mout, vout, vacc, psi1s = ops.mm_fwd(pr, mx, vx, z, ls, sf, A, B)
    psi1, psi2 = go.psi_stats(2 * p['ls'], 2 * p['sf'], p['x'], p['vx'], p['z'])
    m_ref = np.einsum('nm,dm->nd', psi1, p['A'])
    v_ref = np.exp(2 * p['sf']) + np.einsum('dab,nab->nd', p['Bn'], psi2) - m_ref**2
    assert gu.rel_err(N(mout), m_ref) < tol
    assert gu.rel_err(N(vout), v_ref) < tol * 10
    out = ops.mm_bwd(pr, mx, vx, z, ls, sf, A, B, T(p['dm']), T(p['dv']), T(m_ref), vacc, psi1s)
    dm_all = p['dm'] - 2 * p['dv'] * m_ref
    dpsi1 = np.einsum('nd,dm->nm', dm_all, p['A'])
    dpsi2 = np.einsum('nd,dab->nab', p['dv'], p['Bn'])
    r_var, r_l, r_z, r_mu, r_S = go.psi_derivs(dpsi1, psi1, dpsi2, psi2, np.exp(p['ls']),
                                               np.exp(2 * p['sf']), p['x'], p['vx'], p['z'])
    chk = [('dA', np.einsum('nd,nm->dm', dm_all, psi1)), ('dB', np.einsum('nd,nab->dab', p['dv'], psi2)),
           ('dzu', r_z), ('dl', r_l), ('dsf2', np.ravel(r_var)), ('dvsum', np.ravel(p['dv'].sum())),
           ('dmx', r_mu), ('dvx', r_S)]
    for k, ref in chk:
        assert gu.rel_err(N(out[k]), ref) < tol * 10, k


def check_kmat_psi_lik():
    from geepee_b200 import ops
    f = np.load(gu.GOLDEN + '/kernels.npz')
    ls, sf, mx, vx, z = (T(f[k]) for k in ['ls', 'sf', 'mx', 'vx', 'z'])
    assert gu.rel_err(N(ops.kmat(mx, z, ls, sf)), f['kfu']) < 1e-14
    kuu = N(ops.kmat(z, z, ls, sf, jitter=1e-5))
    assert gu.rel_err(kuu, f['Kzz'] + 1e-5 * np.eye(z.shape[0])) < 1e-14
    p1, p2 = ops.psi_stats(mx, vx, z, ls, sf)
    assert gu.rel_err(N(p1), f['psi1']) < 1e-13 and gu.rel_err(N(p2), f['psi2']) < 1e-13
    rng = np.random.RandomState(3)
    m, v, y = rng.standard_normal((50, 3)), rng.rand(50, 3) + 0.1, rng.standard_normal((50, 3))
    sn = np.array([-0.7])
    dm, dv, o = ops.gauss_lik(T(m), T(v), T(y), T(sn), 0.6, -2.5, 0)
    lz, rdm, rdv = go.gauss_log_Z(sn[0], m, v, y, 0.6)
    assert abs(o[0].item() - lz) < 1e-12 * abs(lz) and abs(o[1].item() - rdv.sum()) < 1e-12 * abs(rdv.sum())
    assert gu.rel_err(N(dm), -2.5 * rdm) < 1e-13 and gu.rel_err(N(dv), -2.5 * rdv) < 1e-13
    dm, dv, o = ops.gauss_lik(T(m), T(v), T(y), T(sn), 1.0, -2.5, 1)
    le, rdm, rdv = go.gauss_log_lik_exp(sn[0], m, v, y)
    assert abs(o[0].item() - le) < 1e-12 * abs(le)
    assert abs(-2.5 * o[1].item() - go.gauss_dsn_log_lik_exp(sn[0], m, v, y, -2.5)) < 1e-10
    assert gu.rel_err(N(dm), -2.5 * rdm) < 1e-13 and gu.rel_err(N(dv), -2.5 * rdv) < 1e-13


def check_torch_custom_ops():
    """geepee_b200/torch_ops.py: the library through the PyTorch dispatcher (`torch.ops.geepee_b200.*`) -- golden
    vectors of the reference for kmat / psi_stats (kernels.py:10-22, 181-240), the oracle for the Gaussian likelihood,
    and bit-equality with the direct ctypes wrappers for the moment-matched forward / backward pair."""
    import torch
    import geepee_b200.torch_ops as to
    from geepee_b200 import ops
    g = torch.ops.geepee_b200
    f = np.load(gu.GOLDEN + '/kernels.npz')
    ls, sf, mx, vx, z = (T(f[k]) for k in ['ls', 'sf', 'mx', 'vx', 'z'])
    assert gu.rel_err(N(g.kmat(mx, z, ls, sf)), f['kfu']) < 1e-14
    p1, p2 = g.psi_stats(mx, vx, z, ls, sf)
    assert gu.rel_err(N(p1), f['psi1']) < 1e-13 and gu.rel_err(N(p2), f['psi2']) < 1e-13
    M = z.shape[0]
    kuu = g.kmat(z, z, ls, sf, 1e-5)
    inv, ld = g.spd_inverse(kuu)
    assert gu.rel_err(N(inv), np.linalg.inv(N(kuu))) < 1e-9
    assert abs(ld.item() - np.linalg.slogdet(N(kuu))[1]) < 1e-9 * max(1.0, abs(ld.item()))
    rng = np.random.RandomState(5)
    m, v, y = rng.standard_normal((40, 2)), rng.rand(40, 2) + 0.1, rng.standard_normal((40, 2))
    sn = np.array([-0.4])
    dm, dv, o = g.gauss_lik(T(m), T(v), T(y), T(sn), 0.6, -2.5, 0)
    lz, rdm, rdv = go.gauss_log_Z(sn[0], m, v, y, 0.6)
    assert abs(o[0].item() - lz) < 1e-12 * abs(lz) and gu.rel_err(N(dm), -2.5 * rdm) < 1e-13
    n, Q, Do = mx.shape[0], mx.shape[1], 2
    A = T(rng.standard_normal((Do, M)))
    B = rng.standard_normal((Do, M, M))
    B = T(B + B.transpose(0, 2, 1))
    fw = g.mm_fwd(ops.F64, mx, vx, z, ls, sf, A, B)
    fw_direct = ops.mm_fwd(ops.F64, mx, vx, z, ls, sf, A, B, save=True)
    # outputs reduced with atomics (vacc and what derives from it) may differ in the last bits between two launches
    for a, b in zip(fw, fw_direct):
        assert gu.rel_err(N(a), N(b)) < 1e-13
    # the forward against the materialised statistics: mout = psi1 A^T, vacc = sum_ab B psi2 (aep_models.py:196-198)
    assert gu.rel_err(N(fw[0]), f['psi1'] @ N(A).T) < 1e-12
    assert gu.rel_err(N(fw[2]), np.einsum('nab,dab->nd', f['psi2'], N(B))) < 1e-12
    dmo, dvo = T(rng.standard_normal((n, Do))), T(rng.standard_normal((n, Do)))
    bw = g.mm_bwd(ops.F64, mx, vx, z, ls, sf, A, B, dmo, dvo, fw[0], fw[2], fw[3])
    bw_direct = ops.mm_bwd(ops.F64, mx, vx, z, ls, sf, A, B, dmo, dvo, fw[0], fw[2], fw[3])
    assert len(bw) == len(to.MM_BWD_OUTPUTS)
    for a, k in zip(bw, to.MM_BWD_OUTPUTS):
        assert gu.rel_err(N(a), N(bw_direct[k])) < 1e-12, k
    # dB[d] = sum_n dv[n,d] psi2[n] (aep_models.py:243)
    assert gu.rel_err(N(bw[1]), np.einsum('nd,nab->dab', N(dvo), f['psi2'])) < 1e-12


EMIS_SHAPES = [(37, 3, 2), (300, 4, 4), (129, 5, 3), (64, 8, 8), (50, 1, 6)]


def check_gauss_emis(n, Do, Q):
    """Fused tilted linear-Gaussian emission (lik_layers.py:573-627) through the layer class
    against the oracle's numpy restatement (padded and exact template sizes)."""
    from geepee_b200 import lik_layers
    import torch
    rng = np.random.RandomState(11 * n + Do)
    y = rng.standard_normal((n, Do))
    mx, vx = rng.standard_normal((n, Q)), rng.rand(n, Q) + 0.05
    p = {'C': rng.standard_normal((Do, Q)) * 0.7, 'R': np.log(0.3 + rng.rand(Do)) / 2}
    alpha, scale = 0.7, -3.25
    ref = go.GaussEmis(y, Do, Q)
    ref.set_params(p)
    lz, gi, ge = ref.tilted(mx, vx, alpha, scale, np.arange(n))
    em = lik_layers.Gauss_Emis(y, Do, Q, device=torch.device(DEV))
    em.update_hypers(p)
    lz2, gi2, ge2 = em.compute_emission_tilted(mx, vx, alpha, scale)
    assert abs(lz2 - lz) < 1e-11 * abs(lz)
    for k in ('mx', 'vx'):
        assert gu.rel_err(gi2[k], gi[k]) < 1e-11, k
    for k in ('C', 'R'):
        assert gu.rel_err(ge2[k], ge[k]) < 1e-11, k


def check_probit_lik():
    """probit_lik kernel (lik_layers.py:303-362, 418-436) against the oracle, alpha = 1 (closed
    form), alpha != 1 (Gauss-Hermite) and the VFE expectation."""
    from geepee_b200 import ops
    rng = np.random.RandomState(5)
    m, v = rng.standard_normal((60, 3)) * 1.5, rng.rand(60, 3) * 2 + 0.05
    y = 2.0 * (rng.rand(60, 3) > 0.5) - 1.0
    gx, gw = np.polynomial.hermite.hermgauss(10)
    for alpha in (1.0, 0.5, 0.05):
        dm, dv, o = ops.probit_lik(T(m), T(v), T(y), T(gx), T(gw), alpha, -1.7, 0)
        lz, rdm, rdv = go.probit_log_Z(m, v, y, alpha)
        # device erf / pow differ from scipy's by a few ulp, amplified by pdf**(alpha-1)
        assert abs(o[0].item() - lz) < 1e-10 * abs(lz), alpha
        assert gu.rel_err(N(dm), -1.7 * rdm) < 1e-9 and gu.rel_err(N(dv), -1.7 * rdv) < 1e-9, alpha
    dm, dv, o = ops.probit_lik(T(m), T(v), T(y), T(gx), T(gw), 1.0, 2.5, 1)
    le, rdm, rdv = go.probit_log_lik_exp(m, v, y)
    assert abs(o[0].item() - le) < 1e-10 * abs(le)
    assert gu.rel_err(N(dm), 2.5 * rdm) < 1e-9 and gu.rel_err(N(dv), 2.5 * rdv) < 1e-9


def check_spd_inverse(M, batch):
    """Cluster Gauss-Jordan inverse + log-determinant against numpy (SPD matrices with the
    conditioning of a jittered kernel matrix)."""
    from geepee_b200 import ops
    rng = np.random.RandomState(M + 7 * batch)
    A = np.empty((batch, M, M))
    for b in range(batch):
        z = rng.standard_normal((M, 3))
        d2 = ((z[:, None, :] - z[None, :, :])**2).sum(-1)
        A[b] = np.exp(-0.5 * d2) + 1e-5 * np.eye(M) + (0.1 * b) * np.eye(M)
    inv, ld = ops.spd_inverse(T(A))
    inv, ld = N(inv), N(ld)
    for b in range(batch):
        ref = np.linalg.inv(A[b])
        # residual-based check (the inverse of an ill-conditioned matrix is itself uncertain)
        res = np.abs(inv[b].dot(A[b]) - np.eye(M)).max()
        assert res < 1e-6, (M, b, res)
        assert gu.rel_err(inv[b], ref) < 1e-5, (M, b, gu.rel_err(inv[b], ref))
        sref = np.linalg.slogdet(A[b])[1]
        assert abs(ld[b] - sref) < 1e-8 * max(1.0, abs(sref)), (M, b, ld[b], sref)


def check_tail_primitives(seed=0):
    """Every primitive of the tail program (include/geepee_b200.h GpbTailOp; geepee_b200/tail.py) against
    numpy, with odd sizes, shared (2-D) operands, transposes, strided views and batch sums."""
    from geepee_b200 import tail
    rng = np.random.RandomState(seed)
    r = rng.standard_normal
    # --- GEMM: all four transpose combinations, batch broadcast, C accumulation, both tile shapes
    for (b, m, n, k) in [(3, 37, 29, 41), (1, 5, 7, 3), (2, 64, 64, 64), (5, 70, 130, 33), (40, 96, 96, 20)]:
        for ta in (False, True):
            for tb in (False, True):
                A = r((b,) + ((k, m) if ta else (m, k)))
                B = r(((n, k) if tb else (k, n)))                  # shared operand
                C = r((b, m, n))
                ref = 0.7 * np.einsum('bmk,kn->bmn', np.swapaxes(A, 1, 2) if ta else A, B.T if tb else B) - 1.3 * C
                got = tail.gemm(T(A), T(B), ta=ta, tb=tb, alpha=0.7, C=T(C), beta=-1.3)
                assert gu.rel_err(N(got), ref) < 1e-13, (b, m, n, k, ta, tb)
    A, B = r((4, 20, 9)), r((4, 20, 11))
    got = tail.gemm(T(A).reshape(80, 9), T(B).reshape(80, 11), ta=True, alpha=2.0)     # sum_d A_d^T B_d as ONE product
    assert gu.rel_err(N(got), 2.0 * np.einsum('dka,dkb->ab', A, B)) < 1e-13
    # --- LINCOMB
    S0, S1, S2 = r((3, 6, 6)), r((6, 6)), r((3, 6, 6))
    u, v, u2, v2 = r((3, 6)), r((6,)), r((3, 6)), r((3, 6))
    got = tail.lincomb([(0.5, T(S0)), (-2.0, T(S1), True), (1.5, T(S2), True)], outer=(3.0, T(u), T(v)), eye=0.25)
    ref = 0.5 * S0 - 2.0 * S1.T[None] + 1.5 * np.swapaxes(S2, 1, 2) + 3.0 * u[:, :, None] * v[None, None, :] \
        + 0.25 * np.eye(6)[None]
    assert gu.rel_err(N(got), ref) < 1e-14
    got = tail.lincomb([(1.0, T(S0))], outer=(1.0, T(u), T(u)), outer2=(-2.0, T(u2), T(v2)))
    assert gu.rel_err(N(got), S0 + u[:, :, None] * u[:, None, :] - 2.0 * u2[:, :, None] * v2[:, None, :]) < 1e-14
    got = tail.lincomb([(1.0, T(S0)), (-1.0, T(S2))], reduce=True)
    assert gu.rel_err(N(got), (S0 - S2).sum(0)) < 1e-14
    got = tail.veccomb([(2.0, T(u)), (-1.0, T(u2))])
    assert got.shape == (3, 6) and gu.rel_err(N(got), 2 * u - u2) < 1e-15
    # --- MATVEC
    A0, A1, x0, x1, w0 = r((3, 7, 5)), r((5, 7)), r((3, 5)), r((3, 5)), r((3, 7))
    got = tail.matvec(T(A0), T(x0), c0=0.3, A1=T(A1), x1=T(x1), t1=True, c1=-1.1, w0=T(w0), cw0=2.0)
    ref = 0.3 * np.einsum('bik,bk->bi', A0, x0) - 1.1 * np.einsum('ki,bk->bi', A1, x1) + 2.0 * w0
    assert gu.rel_err(N(got), ref) < 1e-14
    # --- DOTS / total
    a, b2, c = r(1000), r(1000), r((3, 50))
    got = tail.dots([(2.0, T(a), T(b2)), (-1.0, T(c), None), (0.5, T(a), None), (3.0, T(b2), T(b2))], const=0.125)
    ref = 2.0 * a.dot(b2) - c.sum() + 0.5 * a.sum() + 3.0 * b2.dot(b2) + 0.125
    assert abs(got.item() - ref) < 1e-12 * abs(ref)
    big = r(300001)
    assert abs(tail.total(T(big)).item() - big.sum()) < 1e-10
    # --- R packing
    M, Do = 9, 3
    P = M * (M + 1) // 2
    e = 0.3 * r((Do, P))
    R = N(tail.unpack_r(T(e), M))
    iu = np.triu_indices(M)
    Rr = np.zeros((Do, M, M))
    for d in range(Do):
        Rr[d][iu] = e[d]
        Rr[d][np.diag_indices(M)] = np.exp(Rr[d][np.diag_indices(M)])
    assert gu.rel_err(R, Rr) < 1e-15
    dR = r((Do, M, M))
    got = N(tail.pack_r(T(dR), T(Rr), 0.5))
    ref = np.zeros((Do, P))
    for d in range(Do):
        g = dR[d].copy()
        g[np.diag_indices(M)] *= Rr[d][np.diag_indices(M)]
        ref[d] = 0.5 * g[iu]
    assert gu.rel_err(got, ref) < 1e-15
    # --- KHYPER against the reference formulas (golden kernels.npz: d_trace_MKzz_dhypers)
    f = np.load(gu.GOLDEN + '/kernels.npz')
    z, ls, sf, Mm, Kzz = f['z'], f['ls'], f['sf'], f['Mm'], f['Kzz']
    Mz, D = z.shape
    st = np.concatenate([r(Mz * D), r(D), r(1), r(1)])
    out = N(tail.khyper(T(Mm), T(Kzz + 1e-5 * np.eye(Mz)), T(z), T(ls), T(sf), T(st), 1e-5, 0.5))
    dzu0, dl, dsf2, dvsum = st[:Mz * D].reshape(Mz, D), st[Mz * D:Mz * D + D], st[-2], st[-1]
    ref_sf = 2 * np.exp(2 * sf[0]) * (dsf2 + dvsum) + 2 * f['tr_sf']
    ref_ls = dl * np.exp(ls) + 2 * f['tr_ls']
    ref_z = dzu0 + f['tr_z']
    assert abs(out[0] - 0.5 * np.ravel(ref_sf)[0]) < 1e-12 * abs(np.ravel(ref_sf)[0])
    assert gu.rel_err(out[1:1 + D], 0.5 * ref_ls) < 1e-12
    assert gu.rel_err(out[1 + D:].reshape(Mz, D), 0.5 * ref_z) < 1e-12
    # --- GATHER
    parts = [r(5), r((3, 4)), r(1)]
    got = N(tail.gather([T(p) for p in parts], 0.25))
    assert gu.rel_err(got, 0.25 * np.concatenate([p.reshape(-1) for p in parts])) < 1e-15
