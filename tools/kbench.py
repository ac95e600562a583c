#!/usr/bin/env python
"""Per-kernel timing on the B200 (CUDA events on the launching stream, after warm-up).
Writes one JSON line per measurement to gpurun_out/kbench.jsonl.  Development tool."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from geepee_b200 import ops, _lib  # noqa: E402

dev = torch.device('cuda:0')
out_path = os.path.join(os.path.dirname(__file__), '..', 'gpurun_out', 'kbench.jsonl')
os.makedirs(os.path.dirname(out_path), exist_ok=True)
fout = open(out_path, 'a')


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def emit(**kw):
    print(json.dumps(kw))
    fout.write(json.dumps(kw) + '\n')
    fout.flush()


def rnd(*shape, seed=0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).to(dev)


peaks = {}
for name, pr, iters in [('fp64', ops.F64, 20000), ('fp32', ops.F32, 40000)]:
    fl = [0.0]

    def run():
        fl[0] = ops.fma_peak(pr, iters, dev)
    med, mn = timeit(run)
    peaks[name] = fl[0] / (mn * 1e-3) / 1e12
    emit(kind='fma_peak', prec=name, tflops=peaks[name], ms=mn, lib=os.path.basename(_lib.LIB_PATH))

cfgs = [(262144, 256, 10, 1), (262144, 256, 10, 2), (131072, 512, 16, 1), (262144, 128, 5, 1)]
if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    cfgs = cfgs[:1]
if len(sys.argv) > 1 and sys.argv[1] == 'mm':       # pair kernels only (library A/B variants)
    cfgs = []
for n, M, D, Do in cfgs:
    for name, pr in [('fp64', ops.F64), ('fp32', ops.F32)]:
        x, z = rnd(n, D, seed=1), rnd(M, D, seed=2)
        ls, sf = torch.full((D,), 0.5, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
        A = rnd(Do, M, seed=3)
        B = rnd(Do, M, M, seed=4) * 0.01
        B = B + B.transpose(1, 2).contiguous()
        dm, dv = rnd(n, Do, seed=5), rnd(n, Do, seed=6)
        opnd = ops.DetOperands(pr, A, B.contiguous())
        res = {}
        med, mn = timeit(lambda: res.update(r=ops.det_fwd(pr, x, z, ls, sf, opnd, save=True)))
        _, _, Ks, Ts = res['r']
        fl = 2.0 * n * Do * M * M
        emit(kind='det_fwd', prec=name, n=n, M=M, D=D, Do=Do, ms=med, ms_min=mn, tflops=fl / (mn * 1e-3) / 1e12,
             frac_of_fma_peak=fl / (mn * 1e-3) / 1e12 / peaks[name])
        med, mn = timeit(lambda: ops.det_fwd(pr, x, z, ls, sf, opnd, save=False))
        emit(kind='det_fwd_nosave', prec=name, n=n, M=M, D=D, Do=Do, ms=med, ms_min=mn,
             tflops=fl / (mn * 1e-3) / 1e12, frac_of_fma_peak=fl / (mn * 1e-3) / 1e12 / peaks[name])
        med, mn = timeit(lambda: ops.det_bwd(pr, x, z, ls, sf, opnd, dm, dv, Ks, Ts))
        gb = n * opnd.MP * (1 + Do) * (8 if pr == ops.F64 else 4) / 1e9
        emit(kind='det_bwd', prec=name, n=n, M=M, D=D, Do=Do, ms=med, ms_min=mn, gbps=gb / (mn * 1e-3))
        med, mn = timeit(lambda: ops.det_syrk(pr, Ks, dv, M))
        fl = 1.0 * n * Do * M * (M + 1)
        emit(kind='det_syrk', prec=name, n=n, M=M, D=D, Do=Do, ms=med, ms_min=mn, alg_tflops=fl / (mn * 1e-3) / 1e12,
             frac_of_fma_peak=fl / (mn * 1e-3) / 1e12 / peaks[name])
        del Ks, Ts, res

mm_cfgs = [(32768, 256, 2, 2), (32768, 256, 2, 1), (16384, 200, 4, 4), (8192, 128, 5, 4)]
if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    mm_cfgs = mm_cfgs[:1]
for n, M, Q, Do in mm_cfgs:
    for name, pr in ([('fp64', ops.F64)] if sys.argv[1:2] == ['mm'] else [('fp64', ops.F64), ('fp32', ops.F32)]):
        mx, z = rnd(n, Q, seed=1), rnd(M, Q, seed=2)
        vx = (0.1 + torch.rand(n, Q, dtype=torch.float64)).to(dev)
        ls, sf = torch.full((Q,), 0.3, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
        A = rnd(Do, M, seed=3)
        B = (rnd(Do, M, M, seed=4) * 0.01).contiguous()
        dm, dv = rnd(n, Do, seed=5), rnd(n, Do, seed=6)
        res = {}
        med, mn = timeit(lambda: res.update(r=ops.mm_fwd(pr, mx, vx, z, ls, sf, A, B)))
        mout, vacc, psi1s = res['r'][0], res['r'][2], res['r'][3]
        P = M * (M + 1) // 2
        fl_f = n * (P * (4 * Q + 1 + 2 * Do + 2) + M * (6 * Q + 2 * Do))
        emit(kind='mm_fwd', prec=name, n=n, M=M, Q=Q, Do=Do, ms=med, ms_min=mn, rows_per_s=n / (mn * 1e-3),
             alg_tflops=fl_f / (mn * 1e-3) / 1e12)
        med, mn = timeit(lambda: ops.mm_bwd(pr, mx, vx, z, ls, sf, A, B, dm, dv, mout, vacc, psi1s))
        fl_b = n * (P * (18 * Q + 6 * Do + 6) + M * (14 * Q + 6 * Do)) - fl_f
        emit(kind='mm_bwd', prec=name, n=n, M=M, Q=Q, Do=Do, ms=med, ms_min=mn, rows_per_s=n / (mn * 1e-3),
             alg_tflops=fl_b / (mn * 1e-3) / 1e12)
print('done')
