"""Host side of the replicated O(Dout M^3) tail: thin tensor wrappers over the library's tail
primitives (include/geepee_b200.h: GpbTailOp / gpb_tail_exec / gpb_tail_gather; kernels in
csrc/gpb_tail.cuh).

The reference writes q(u) / cavity / log-partition algebra and the chain rules from the reduced
statistics back to the parameters as numpy einsum + linalg calls (base_models.py:454-516,630-658,
aep_models.py:62-114,252-297,462-586, vfe_models.py:309-325,363-394,518-541, kernels.py:447-475).
layers.py states the same algebra with the calls below; every one of them is ONE launch of a
library kernel on torch's current stream (fp64, batched over the output dimensions, FP64
tensor cores for the products).  torch supplies device memory only.

Operands are fp64 device tensors whose last dimension is contiguous; a 2-D operand is shared by
the whole batch (batch stride 0), so `Kuuinv` is never replicated.
"""
import ctypes

import torch

from . import _lib, ops

GEMM, LINCOMB, MATVEC, DOTS, UNPACK_R, PACK_R, KHYPER, SUM = range(1, 9)


class GpbTailOp(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int), ('flags', ctypes.c_int), ('batch', ctypes.c_int), ('m', ctypes.c_int),
                ('n', ctypes.c_int), ('k', ctypes.c_int), ('src', ctypes.c_void_p * 6),
                ('sstride', ctypes.c_long * 6), ('ld', ctypes.c_int * 6), ('coef', ctypes.c_double * 8),
                ('dst', ctypes.c_void_p), ('dstride', ctypes.c_long), ('ldd', ctypes.c_int)]


_F = torch.float64


def _check(t):
    if t.device.type != _lib.device_type():
        raise RuntimeError('geepee_b200: tensor on %s, library expects %s (no CPU fallback)'
                           % (t.device.type, _lib.device_type()))
    if t.dtype != _F or (t.dim() > 0 and t.numel() > 1 and t.stride(-1) != 1):
        raise RuntimeError('geepee_b200.tail: expected fp64 tensor with a contiguous last dimension')
    return t


def _mat(t):
    """-> (tensor, batch or None, rows, cols, ld, batch stride) of a [rows, cols] / [b, rows, cols] operand."""
    _check(t)
    if t.dim() == 2:
        return t, None, t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else t.shape[1]), 0
    if t.dim() == 3:
        ld = t.stride(1) if t.shape[1] > 1 else t.shape[2]
        return t, t.shape[0], t.shape[1], t.shape[2], ld, (t.stride(0) if t.shape[0] > 1 else 0)
    raise RuntimeError('geepee_b200.tail: matrix operand must be 2-D or 3-D, got %d-D' % t.dim())


def _vec(t):
    """-> (tensor, batch or None, length, batch stride) of a [len] / [b, len] operand."""
    _check(t)
    if t.dim() == 1:
        return t, None, t.shape[0], 0
    if t.dim() == 2:
        return t, t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else 0)
    raise RuntimeError('geepee_b200.tail: vector operand must be 1-D or 2-D, got %d-D' % t.dim())


def _batch(*bs):
    b = None
    for x in bs:
        if x is None:
            continue
        if b is None or b == 1:
            b = x
        elif x not in (1, b):
            raise RuntimeError('geepee_b200.tail: batch sizes %r do not broadcast' % (bs,))
    return b


def _run(op, like):
    rc = _lib.get().gpb_tail_exec(ctypes.byref(op), 1, ops._stream(like))
    ops._chk(rc, 'tail_exec')


def _set(op, i, t, stride, ld):
    op.src[i] = t.data_ptr()
    op.sstride[i] = int(stride)
    op.ld[i] = int(ld)


def gemm(A, B, ta=False, tb=False, alpha=1.0, C=None, beta=0.0, out=None):
    """out[b] = alpha op(A[b]) op(B[b]) + beta C[b]   (FP64 tensor cores; 2-D operands are shared)."""
    A, ba, ra, ca, lda, sa = _mat(A)
    B, bb, rb, cb, ldb, sb = _mat(B)
    m, k = (ca, ra) if ta else (ra, ca)
    k2, n = (cb, rb) if tb else (rb, cb)
    if k != k2:
        raise RuntimeError('geepee_b200.tail.gemm: inner dimensions %d and %d differ' % (k, k2))
    bc = None
    if C is not None:
        C, bc, rc_, cc, ldc, sc = _mat(C)
        if (rc_, cc) != (m, n):
            raise RuntimeError('geepee_b200.tail.gemm: C has shape %r, expected %r' % ((rc_, cc), (m, n)))
    b = _batch(ba, bb, bc)
    if out is None:
        out = torch.empty((m, n) if b is None else (b, m, n), dtype=_F, device=A.device)
    O, bo, ro, co, ldo, so = _mat(out)
    op = GpbTailOp()
    op.kind, op.flags = GEMM, (1 if ta else 0) | (2 if tb else 0)
    op.batch, op.m, op.n, op.k = (b or 1), m, n, k
    _set(op, 0, A, sa, lda)
    _set(op, 1, B, sb, ldb)
    if C is not None:
        _set(op, 2, C, sc, ldc)
    op.coef[0], op.coef[1] = float(alpha), float(beta)
    op.dst, op.dstride, op.ldd = O.data_ptr(), int(so if bo else 0), int(ldo)
    _run(op, A)
    return out


def lincomb(terms, outer=None, outer2=None, eye=0.0, reduce=False, out=None):
    """out[b] = sum_s c_s op(S_s[b]) + c u[b] v[b]^T (+ c' u'[b] v'[b]^T) + eye I.
    terms: up to four (coef, matrix[, transposed]) -- two when `outer2` is given; outer / outer2:
    (coef, u, v).  reduce=True sums the batch into one matrix (a shared operand then counts once per
    batch element).  Vectors are handled as [b, 1, n] matrices."""
    if len(terms) > (2 if outer2 is not None else 4):
        raise RuntimeError('geepee_b200.tail.lincomb: too many terms')
    op = GpbTailOp()
    op.kind = LINCOMB
    flags, bs, shape, dev = 0, [], None, None
    for s, term in enumerate(terms):
        c, S = term[0], term[1]
        tr = len(term) > 2 and term[2]
        if S.dim() == 1:
            S = S.reshape(1, -1)
        S, b_, r, cdim, ld, st = _mat(S)
        if tr:
            r, cdim = cdim, r
            flags |= 1 << s
        if shape is None:
            shape = (r, cdim)
        elif shape != (r, cdim):
            raise RuntimeError('geepee_b200.tail.lincomb: term shapes %r and %r differ' % (shape, (r, cdim)))
        bs.append(b_)
        _set(op, s, S, st, ld)
        op.coef[s] = float(c)
        dev = S
    for slot, o in ((4, outer), (2, outer2)):
        if o is None:
            continue
        c, u, v = o
        u, bu, lu, su = _vec(u)
        v, bv, lv, sv = _vec(v)
        if shape is None:
            shape = (lu, lv)
        elif shape != (lu, lv):
            raise RuntimeError('geepee_b200.tail.lincomb: outer product shape %r, expected %r' % ((lu, lv), shape))
        bs += [bu, bv]
        _set(op, slot, u, su, 1)
        _set(op, slot + 1, v, sv, 1)
        op.coef[4 if slot == 4 else 2] = float(c)
        if slot == 2:
            flags |= 1 << 9
        dev = u
    b = _batch(*bs)
    if reduce:
        flags |= 1 << 8
    op.flags = flags
    op.batch, op.m, op.n = (b or 1), shape[0], shape[1]
    op.coef[5] = float(eye)
    if out is None:
        out = torch.empty(shape if (b is None or reduce) else (b,) + shape, dtype=_F, device=dev.device)
    O = out
    if O.dim() == 1:
        O = O.reshape(1, -1)
    O, bo, ro, co, ldo, so = _mat(O)
    op.dst, op.dstride, op.ldd = O.data_ptr(), int(so if bo else 0), int(ldo)
    _run(op, dev)
    return out


def veccomb(terms, out=None):
    """out[b, :] = sum_s c_s x_s[b, :]  (vectors [len] or [b, len])."""
    mats = [(c, x.reshape(1, -1) if x.dim() == 1 else x.unsqueeze(-2)) for c, x in terms]
    r = lincomb(mats, out=None if out is None else (out.reshape(1, -1) if out.dim() == 1 else out.unsqueeze(-2)))
    if out is not None:
        return out
    return r.reshape(-1) if all(x.dim() == 1 for _, x in terms) else r.squeeze(-2)


def matvec(A0, x0, t0=False, c0=1.0, A1=None, x1=None, t1=False, c1=1.0, w0=None, cw0=1.0, w1=None, cw1=1.0,
           out=None):
    """out[b] = c0 op(A0[b]) x0[b] + c1 op(A1[b]) x1[b] + cw0 w0[b] + cw1 w1[b]."""
    op = GpbTailOp()
    op.kind = MATVEC
    A0, ba, r0, c0_, lda, sa = _mat(A0)
    m, k = (c0_, r0) if t0 else (r0, c0_)
    x0, bx, lx, sx = _vec(x0)
    if lx != k:
        raise RuntimeError('geepee_b200.tail.matvec: vector length %d, expected %d' % (lx, k))
    _set(op, 0, A0, sa, lda)
    _set(op, 1, x0, sx, 1)
    op.coef[0] = float(c0)
    bs = [ba, bx]
    flags = 1 if t0 else 0
    if A1 is not None:
        A1, ba1, r1, c1_, lda1, sa1 = _mat(A1)
        m1, k1 = (c1_, r1) if t1 else (r1, c1_)
        x1, bx1, lx1, sx1 = _vec(x1)
        if (m1, k1) != (m, k) or lx1 != k:
            raise RuntimeError('geepee_b200.tail.matvec: second product has a different shape')
        _set(op, 2, A1, sa1, lda1)
        _set(op, 3, x1, sx1, 1)
        op.coef[1] = float(c1)
        bs += [ba1, bx1]
        flags |= 2 if t1 else 0
    for slot, w, cw, ci in ((4, w0, cw0, 2), (5, w1, cw1, 3)):
        if w is not None:
            w, bw, lw, sw = _vec(w)
            if lw != m:
                raise RuntimeError('geepee_b200.tail.matvec: addend length %d, expected %d' % (lw, m))
            _set(op, slot, w, sw, 1)
            op.coef[ci] = float(cw)
            bs.append(bw)
    b = _batch(*bs)
    op.flags = flags
    op.batch, op.m, op.k = (b or 1), m, k
    if out is None:
        out = torch.empty((m,) if b is None else (b, m), dtype=_F, device=A0.device)
    O, bo, lo, so = _vec(out)
    op.dst, op.dstride = O.data_ptr(), int(so if bo else 0)
    _run(op, A0)
    return out


def dots(terms, const=0.0, out=None, acc=False):
    """out[0] = (acc ? out[0] : 0) + const + sum_t c_t <x_t, y_t>   (y_t None: sum of x_t); contiguous operands."""
    dev = terms[0][1]
    if out is None:
        out = torch.empty(1, dtype=_F, device=dev.device)
    for i0 in range(0, len(terms), 3):
        op = GpbTailOp()
        op.kind, op.flags = DOTS, (1 if (acc or i0 > 0) else 0)
        op.batch = 1
        for t, (c, x, y) in enumerate(terms[i0:i0 + 3]):
            _check(x)
            if not x.is_contiguous() or (y is not None and (not y.is_contiguous() or y.numel() != x.numel())):
                raise RuntimeError('geepee_b200.tail.dots: operands must be contiguous and of equal size')
            op.src[2 * t] = x.data_ptr()
            op.ld[2 * t] = x.numel()
            if y is not None:
                op.src[2 * t + 1] = _check(y).data_ptr()
            op.coef[t] = float(c)
        op.coef[6] = float(const) if i0 == 0 else 0.0
        op.dst = out.data_ptr()
        _run(op, dev)
    return out


def total(x, out=None):
    """out[0] = sum of all elements of a (large) contiguous tensor; two launches."""
    _check(x)
    if not x.is_contiguous():
        raise RuntimeError('geepee_b200.tail.total: operand must be contiguous')
    if out is None:
        out = torch.empty(1, dtype=_F, device=x.device)
    scratch = torch.empty(1024, dtype=_F, device=x.device)
    op = GpbTailOp()
    op.kind, op.batch = SUM, 1
    op.src[0], op.sstride[0] = x.data_ptr(), x.numel()
    op.src[1] = scratch.data_ptr()
    op.dst = out.data_ptr()
    _run(op, x)
    return out


def unpack_r(eta1, M):
    """base_models.py:645-653: eta1_R[Dout, M(M+1)/2] -> upper-triangular R[Dout, M, M], diagonal exponentiated."""
    _check(eta1)
    Do = eta1.shape[0]
    out = torch.empty((Do, M, M), dtype=_F, device=eta1.device)
    op = GpbTailOp()
    op.kind, op.batch, op.m = UNPACK_R, Do, M
    _set(op, 0, eta1, eta1.stride(0) if Do > 1 else 0, 1)
    op.dst, op.dstride, op.ldd = out.data_ptr(), M * M, M
    _run(op, eta1)
    return out


def pack_r(dR, R, coef=1.0, out=None):
    """base_models.py:505-514: triu(dR) with the diagonal times diag(R) -> [Dout, M(M+1)/2]."""
    dR, bd, M, _, ldd, sd = _mat(dR)
    R, br, _, _, ldr, sr = _mat(R)
    Do = bd or 1
    P = M * (M + 1) // 2
    if out is None:
        out = torch.empty((Do, P), dtype=_F, device=dR.device)
    op = GpbTailOp()
    op.kind, op.batch, op.m = PACK_R, Do, M
    _set(op, 0, dR, sd, ldd)
    _set(op, 1, R, sr, ldr)
    op.coef[0] = float(coef)
    op.dst, op.dstride = out.data_ptr(), P
    _run(op, dR)
    return out


def khyper(Mm, Kuu, zu, ls, sf, stats, jitter, scale=1.0, out=None):
    """kernels.py:447-475 (d_trace_MKzz_dhypers with Kzz = Kuu - jitter I) folded with the direct kernel
    derivatives (aep_models.py:455-460,497-504).  stats: contiguous [dzu0[M*D] | dl[D] | dsf2 | dvsum].
    -> record [dsf | dls[D] | dzu[M*D]] * scale."""
    M, D = zu.shape
    for t in (Mm, Kuu, zu, ls, sf, stats):
        _check(t)
    if stats.numel() != M * D + D + 2 or not stats.is_contiguous():
        raise RuntimeError('geepee_b200.tail.khyper: statistics record has the wrong size')
    need = int(_lib.get().gpb_tail_khyper_out_len(M, D))       # record + per-block scratch
    if out is None:
        out = torch.empty(need, dtype=_F, device=zu.device)
    elif out.numel() < need:
        raise RuntimeError('geepee_b200.tail.khyper: output buffer too small (%d < %d)' % (out.numel(), need))
    op = GpbTailOp()
    op.kind, op.batch, op.m, op.k = KHYPER, 1, M, D
    _set(op, 0, Mm, 0, Mm.stride(0))
    _set(op, 1, Kuu, 0, Kuu.stride(0))
    _set(op, 2, zu, 0, D)
    _set(op, 3, ls, 0, 1)
    _set(op, 4, sf, 0, 1)
    _set(op, 5, stats, 0, 1)
    op.coef[0], op.coef[1] = float(jitter), float(scale)
    op.dst = out.data_ptr()
    _run(op, zu)
    return out[:1 + D + M * D]


def multicopy(dsts, srcs):
    """dst_i <- src_i for contiguous fp64 tensors of equal sizes, ONE launch (per 24 pairs)."""
    n = len(dsts)
    if n == 0:
        return
    for d, s_ in zip(dsts, srcs):
        _check(d)
        _check(s_)
        if not (d.is_contiguous() and s_.is_contiguous()) or d.numel() != s_.numel():
            raise RuntimeError('geepee_b200.tail.multicopy: tensors must be contiguous and of equal size')
    sp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    dp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in dsts])
    counts = (ctypes.c_long * n)(*[t.numel() for t in srcs])
    desc = (n, sp, dp, counts)
    multicopy_prepared(desc, dsts[0])
    return desc


def multicopy_prepared(desc, like):
    """Re-issue a multicopy whose tensors (addresses, sizes) are unchanged: `desc` is what multicopy returned."""
    n, sp, dp, counts = desc
    rc = _lib.get().gpb_tail_copy(n, sp, dp, counts, ops._stream(like))
    ops._chk(rc, 'tail_copy')


def gather(tensors, scale=1.0, out=None):
    """out = scale * concat(t.reshape(-1) for t in tensors): one launch (per 40 sources)."""
    n = len(tensors)
    for t in tensors:
        _check(t)
        if not t.is_contiguous():
            raise RuntimeError('geepee_b200.tail.gather: sources must be contiguous')
    tot = sum(t.numel() for t in tensors)
    if out is None:
        out = torch.empty(tot, dtype=_F, device=tensors[0].device)
    srcs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors])
    counts = (ctypes.c_long * n)(*[t.numel() for t in tensors])
    rc = _lib.get().gpb_tail_gather(n, srcs, counts, float(scale), ctypes.c_void_p(out.data_ptr()),
                                    ops._stream(out))
    ops._chk(rc, 'tail_gather')
    return out
