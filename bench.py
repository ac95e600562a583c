#!/usr/bin/env python
"""bench.py -- AEP energy+gradient data-points/sec of the geepee hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--prec fp64|fp32]
    python bench.py --impl reference ...        # the reference's CPU implementation, same metric
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full `objective_function` call (energy + every gradient) over one batch of
synthetic data of the configured shape.  `value` is timed with the data resident in HBM;
`e2e` is the same call through the public model API with the step's inputs copied from pinned
host memory and the gradients read back, inside the timed region.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ------------------------------------------------------------------------------------------
WORKLOADS = {
    # BASELINE.json configs[2] -- the shape the north-star target is quoted on (N=1M, D=10, M=256)
    'cfg3_sdgpr': dict(model='SDGPR', N=1000000, D=10, hidden=[2, 2], Do=1, M=256, alpha=1.0, seed=3,
                       desc='aep.SDGPR 2 hidden layers [2,2], N=1e6, D_in=10, M=256/layer, alpha=1.0, '
                            'moment propagation, full batch'),
    'ns_sgpr': dict(model='SGPR', N=1000000, D=10, Do=1, M=256, alpha=0.5, seed=5,
                    desc='aep.SGPR N=1e6, D=10, M=256, alpha=0.5, full batch'),
    'cfg5_sgpr': dict(model='SGPR', N=10000000, D=16, Do=1, M=512, alpha=0.5, seed=5,
                      desc='aep.SGPR N=1e7, D=16, M=512, alpha=0.5, full batch'),
    # BASELINE.json configs[1] and configs[3] (parity-test shapes; measurable with --workload)
    'cfg2_sgplvm': dict(model='SGPLVM', N=100000, Q=5, Do=50, M=128, alpha=0.5, seed=2,
                        desc='aep.SGPLVM N=1e5, D_out=50, latent Q=5, M=128, alpha=0.5, full batch'),
    'cfg4_sgpssm': dict(model='SGPSSM', N=1000000, Q=4, Do=4, M=200, alpha=0.5, seed=4,
                        desc='aep.SGPSSM T=1e6, latent dim 4, M=200, alpha=0.5, linear-Gaussian emission, '
                             'full window'),
    'cfg1_sgpr': dict(model='SGPR', N=200, D=1, Do=1, M=50, alpha=0.5, seed=42,
                      desc='aep.SGPR N=200, D=1, M=50, alpha=0.5 (examples/gpr_aep_examples.py)'),
    'small_sdgpr': dict(model='SDGPR', N=20000, D=10, hidden=[2, 2], Do=1, M=64, alpha=1.0, seed=3,
                        desc='aep.SDGPR [2,2], N=2e4, D_in=10, M=64 (smoke-sized)'),
}


def f_det(D, M, Do):
    return 8 * D * M + Do * (2 * M * M + M * (M + 1) + 10 * M)


def f_mm(Q, M, Do):
    P = M * (M + 1) // 2
    return P * (18 * Q + 6 * Do + 6) + M * (14 * Q + 6 * Do)


def flops_per_row(w):
    """Algorithmic flops per data point (SURVEY.md section 8d formulas)."""
    if w['model'] == 'SGPR':
        return f_det(w['D'], w['M'], w['Do'])
    if w['model'] == 'SGPLVM':
        return f_mm(w['Q'], w['M'], w['Do'])
    if w['model'] == 'SGPSSM':
        return f_mm(w['Q'], w['M'], w['Q'])      # dynamics layer; emission + x terms add < 1 %
    sizes = [w['D']] + list(w['hidden']) + [w['Do']]
    tot = f_det(sizes[0], w['M'], sizes[1])
    for i in range(1, len(sizes) - 1):
        tot += f_mm(sizes[i], w['M'], sizes[i + 1])
    return tot


def make_data(w, n=None):
    """Synthetic data of the workload's shape (fixed seed)."""
    n = w['N'] if n is None else n
    rng = np.random.RandomState(w['seed'])
    if w['model'] == 'SGPLVM':     # SURVEY.md 8d cfg 2: Y = tanh(X* W1) W2 + 0.1 noise, standardised
        Xs = rng.standard_normal((n, w['Q']))
        W1, W2 = rng.standard_normal((w['Q'], 20)), rng.standard_normal((20, w['Do']))
        Y = np.tanh(Xs.dot(W1)).dot(W2) + 0.1 * rng.standard_normal((n, w['Do']))
        Y = (Y - Y.mean(0)) / Y.std(0)
        return Xs, Y
    if w['model'] == 'SGPSSM':     # SURVEY.md 8d cfg 4 (fallback): 4-D damped nonlinear oscillator
        t = np.arange(n) * 0.05
        ph = rng.uniform(0, 2 * np.pi, w['Do'])
        fr = np.array([1.0, 1.7, 0.6, 2.3])[:w['Do']]
        Y = np.stack([np.sin(fr[i] * t + ph[i]) * (1.0 + 0.3 * np.cos(0.11 * fr[i] * t)) for i in range(w['Do'])], 1)
        Y = np.tanh(1.5 * Y) + 0.05 * rng.standard_normal((n, w['Do']))
        Y = (Y - Y.mean(0)) / Y.std(0)
        return None, Y
    if w['model'] == 'SGPR' and w['D'] == 1:
        X = rng.rand(n, 1)
        Y = np.sin(12 * X) + 0.5 * np.cos(25 * X) + rng.randn(n, 1) * 0.2
        return X, Y
    X = rng.standard_normal((n, w['D']))
    wv = rng.standard_normal((w['D'], w['Do'])) / np.sqrt(w['D'])
    Y = np.sin(X.dot(wv)) + 0.1 * rng.standard_normal((n, w['Do']))
    return X, Y


def _layer_recipe(N, M, Din, Dout, X, sfx=''):
    """Base_SGP_Layer.init_hypers recipe (base_models.py:518-597) without building a model."""
    from geepee_b200 import layers
    lay = layers.Base_SGP_Layer.__new__(layers.Base_SGP_Layer)
    lay.N, lay.M, lay.Din, lay.Dout, lay.nat_param = N, M, Din, Dout, True
    return layers.Base_SGP_Layer.init_hypers(lay, X, key_suffix=sfx)


def make_params(model, Y, w=None, X=None):
    np.random.seed(0)
    if w is not None and w['model'] == 'SGPLVM':
        # hand-built (SURVEY.md 8d cfg 2): skips the nested GPR fit of base_models.py:839-881
        p = _layer_recipe(Y.shape[0], w['M'], w['Q'], w['Do'], X)
        p['ls'] = np.zeros(w['Q'])
        p['sf'] = np.zeros(1)
        p['sn'] = np.array(np.log(0.1))
        p['x1'] = X / 0.1
        p['x2'] = 0.5 * np.log(1.0 / 0.1 - 1.0) * np.ones_like(X)
        return p
    if w is not None and w['model'] == 'SGPSSM':
        # SURVEY.md 8d cfg 4: x factors from base_models.py:1628-1634, C = I, R / sn as in
        # examples/gpssm_hodgkin_huxley.py:328-329, dynamics layer from the layer recipe
        p = _layer_recipe(Y.shape[0] - 1, w['M'], w['Q'], w['Q'], Y[:-1], '_dynamic')
        p['x_factor_1'] = Y / 0.1 / 3.0
        p['x_factor_2'] = 0.5 * np.log(10.0 / 3.0) * np.ones_like(Y)
        p['C_emission'] = np.eye(w['Q'])
        p['R_emission'] = np.log(0.01) / 2 * np.ones(w['Do'])
        p['sn'] = np.log(0.01) / 2 * np.ones(1)
        return p
    p = model.init_hypers(Y)
    p['sn'] = np.array(np.log(0.1))
    return p


# ------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref if it was built) or the oracle port
# ------------------------------------------------------------------------------------------
def cpu_model_factory():
    """-> (kind, build(w, X, Y) -> model with objective_function/init_hypers)."""
    ref_pkg = os.path.join(ROOT, 'oracle', '_ref')
    if os.path.exists(os.path.join(ref_pkg, 'geepee', 'aep_models.py')):
        try:
            sys.path.insert(0, ref_pkg)
            sys.path.append(os.path.join(ref_pkg, 'stubs'))
            import importlib
            aep = importlib.import_module('geepee.aep_models')

            def build(w, X, Y):
                if w['model'] == 'SGPR':
                    return aep.SGPR(X, Y, w['M'], lik='Gaussian')
                if w['model'] == 'SGPLVM':
                    return aep.SGPLVM(Y, w['Q'], w['M'], lik='Gaussian')
                if w['model'] == 'SGPSSM':
                    return aep.SGPSSM(Y, w['Q'], w['M'], lik='Gaussian')
                return aep.SDGPR(X, Y, w['M'], w['hidden'], lik='Gaussian')
            return 'reference', build
        except Exception:  # noqa: BLE001
            pass
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import geepee_oracle as go

    def build(w, X, Y):
        if w['model'] == 'SGPR':
            return go.AepSGPR(X, Y, w['M'])
        if w['model'] == 'SGPLVM':
            return go.AepSGPLVM(Y, w['Q'], w['M'])
        if w['model'] == 'SGPSSM':
            return go.AepSGPSSM(Y, w['Q'], w['M'])
        return go.AepSDGPR(X, Y, w['M'], w['hidden'])
    return 'port', build


def cpu_params(w, X, Y):
    """Same recipe as the GPU arm (geepee_b200 layers' init_hypers is the reference's)."""
    import io
    import contextlib
    from geepee_b200 import layers
    np.random.seed(0)
    if w['model'] in ('SGPLVM', 'SGPSSM'):
        with contextlib.redirect_stdout(io.StringIO()):
            return make_params(None, Y, w, X)
    sizes = [w['D']] + list(w.get('hidden', [])) + [w['Do']]
    p = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(len(sizes) - 1):
            lay = layers.Base_SGP_Layer.__new__(layers.Base_SGP_Layer)
            lay.N, lay.M, lay.Din, lay.Dout, lay.nat_param = X.shape[0], w['M'], sizes[i], sizes[i + 1], True
            sfx = '' if w['model'] == 'SGPR' else '_%d' % i
            Xi = None
            if i == 0:
                Xi = X
                if X.shape[0] < 2 * w['M']:      # tiny calibration samples: kmeans needs >= M points
                    Xi = np.vstack([X, np.random.standard_normal((2 * w['M'] - X.shape[0], X.shape[1]))])
                lay.N = Xi.shape[0]
            p.update(layers.Base_SGP_Layer.init_hypers(lay, Xi, key_suffix=sfx))
    p['sn'] = np.array(np.log(0.1))
    return p


def time_cpu(w, budget_s, steps=1, warmup=0):
    """Throughput of the CPU implementation on a bounded sample of the workload.

    One reference call costs T(n) = a + b*n: `a` is the data-independent O(Dout M^4) tail (the
    reference's three-operand einsums, e.g. aep_models.py:252,258: 68 s at M=256 on this class
    of host) and `b` the per-row cost.  At the workload's N the tail is amortised, so the
    rows/s the reference would sustain on the full workload is 1/b.  Two calls (n1 < n2) give
    a and b; the number reported is 1/b, with a stated alongside.  `steps`/`warmup` beyond one
    measurement pair are not repeated: a single pair already takes minutes at M=256."""
    import copy
    kind, build = cpu_model_factory()

    def call(n):
        X, Y = make_data(w, n)
        m = build(w, X, Y)
        p = cpu_params(w, X, Y)
        t = time.perf_counter()
        m.objective_function(copy.deepcopy(p), n, alpha=w['alpha'])
        return time.perf_counter() - t

    n1 = min(64 if w['model'] in ('SGPR', 'SDGPR') else 2 * w['M'], w['N'])
    t1 = call(n1)
    if n1 >= w['N']:
        return dict(value=n1 / t1, unit='rows/s', cores=1, threads_available=os.cpu_count(), kind=kind,
                    sample='full workload (%d rows), one call, %.3f s' % (n1, t1), ms_per_step=t1 * 1e3, rows=n1)
    # size the second sample from a quick per-row probe so that it adds about budget_s
    n_probe = min(4 * n1, w['N'])
    t_probe = call(n_probe) if t1 < 5.0 else None
    if t_probe is not None and t_probe > t1:
        b0 = (t_probe - t1) / (n_probe - n1)
        n2 = int(max(4 * n1, min(budget_s / b0, 20000, w['N'])))
    else:
        n2 = int(min(1024, w['N']))
    t2 = call(n2)
    b = (t2 - t1) / (n2 - n1)
    if b <= 0:
        b = t2 / n2
    a = max(t1 - b * n1, 0.0)
    return dict(value=1.0 / b, unit='rows/s', cores=1, threads_available=os.cpu_count(), kind=kind,
                sample='two calls of the %s shape (M=%d): n=%d in %.2f s, n=%d in %.2f s -> per-row %.3f ms '
                       '(reported as rows/s), data-independent tail %.1f s per call; numpy einsum contractions '
                       'are single-threaded, BLAS parts use default threading'
                       % (w['model'], w['M'], n1, t1, n2, t2, b * 1e3, a),
                ms_per_step=t2 * 1e3, rows=n2, tail_s=a, per_row_ms=b * 1e3)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], [], set(), []
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], c[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


# ------------------------------------------------------------------------------------------
def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = time_cpu(w, budget_s=20.0)
    line = {
        'impl': 'reference', 'metric': 'AEP energy+grad data-points/sec', 'value': r['value'], 'unit': 'rows/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload + ': ' + w['desc'], 'rows_per_step': r['rows']},
        'cpu_baseline': {'value': r['value'], 'unit': 'rows/s', 'cores': r['cores'],
                         'threads_available': r['threads_available'], 'kind': r['kind'], 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': 'rows/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_gpu(args, w):
    import torch
    import torch.distributed as tdist
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        tdist.init_process_group('nccl', device_id=dev)
    from geepee_b200 import aep_models as aep, ops

    X, Y = make_data(w)
    N = w['N']
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        if w['model'] == 'SGPR':
            model = aep.SGPR(X, Y, w['M'], prec=args.prec, device=dev)
        elif w['model'] == 'SGPLVM':
            model = aep.SGPLVM(Y, w['Q'], w['M'], prec=args.prec, device=dev)
        elif w['model'] == 'SGPSSM':
            model = aep.SGPSSM(Y, w['Q'], w['M'], prec=args.prec, device=dev)
        else:
            model = aep.SDGPR(X, Y, w['M'], w['hidden'], prec=args.prec, device=dev)
        params = make_params(model, Y, w, X)
    alpha = w['alpha']

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def step():
        return model.objective_function(params, N, alpha=alpha)

    # pinned host copies of the step's inputs for the end-to-end leg
    xh = torch.from_numpy(X).pin_memory() if w['model'] in ('SGPR', 'SDGPR') else None
    yh = torch.from_numpy(Y).pin_memory()

    # each rank's inputs for a step are its own contiguous slice of the rows (geepee_b200/dist.py)
    lo, hi = (rank * N) // world, ((rank + 1) * N) // world

    def step_e2e():
        if xh is not None:
            model._x[lo:hi].copy_(xh[lo:hi], non_blocking=True)
        model._y[lo:hi].copy_(yh[lo:hi], non_blocking=True)
        return model.objective_function(params, N, alpha=alpha)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(max(args.warmup, 3)):
        energy, grads = step()
    # ---- device-resident timing, with per-kernel events and clock sampling ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ops.profile_enable(True)
    ops.profile_collect()
    l0 = ops.launch_count()
    ms_total, (energy, grads) = timed(step, args.steps)
    launches = ops.launch_count() - l0
    prof = ops.profile_collect()
    ops.profile_enable(False)
    clocks = sampler.stop() if sampler else {}
    ms_step = ms_total / args.steps
    value = N / (ms_step * 1e-3)
    # ---- end-to-end timing ----
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    ms_e2e /= args.steps
    h2d = ((X[lo:hi].nbytes if xh is not None else 0) + Y[lo:hi].nbytes) * world + world * sum(np.asarray(v).nbytes for v in params.values())
    d2h = world * (8 + sum(np.asarray(v).nbytes for v in grads.values()))
    # ---- FMA-pipe peak of this box (the binding roofline is the FP64 / FP32 FMA pipe) ----
    pr = ops.PREC[args.prec]
    flops_box = [0.0]

    def peak_run():
        flops_box[0] = ops.fma_peak(pr, 20000 if pr == ops.F64 else 40000, dev)
    peak_run()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        peak_run()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peak_tf = flops_box[0] / (best * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return
    # dominant kernel + its algorithmic flops per launch, rows of ONE rank
    rows_rank = N // world
    if w['model'] in ('SDGPR', 'SGPLVM', 'SGPSSM'):
        if w['model'] == 'SDGPR':
            sizes = [w['D']] + list(w['hidden']) + [w['Do']]
        else:
            sizes = [0, w['Q'], w['Do'] if w['model'] == 'SGPLVM' else w['Q']]
        P = w['M'] * (w['M'] + 1) // 2
        fl = sum(rows_rank * P * (14 * sizes[i] + 4 * sizes[i + 1] + 5) for i in range(1, len(sizes) - 1))
        slot, kname = 'mm_pairs_bwd', 'mm_pairs_kernel<T,Q,DOC,BWD=true> (psi2 regenerated on chip; all moment-matched layers)'
        if pr == ops.F64 and sizes[-1] > 4 and sizes[1] <= 8:
            kname = ('mm_bwd_wide_mma_kernel<Q> (psi2 once per row and pair on chip, the four backward '
                     'contractions as DMMA.8x8x4 tiles on the FP64 tensor cores)')
    else:
        fl = rows_rank * 2.0 * w['Do'] * w['M'] ** 2
        slot = 'det_fwd'
        kname = ('det_fwd_mma_kernel<MP> (Kfu generation fused with Kfu.B on the FP64 tensor cores, DMMA.8x8x4)'
                 if pr == ops.F64 else 'det_fwd_kernel<float,MP> (Kfu generation fused with Kfu.B, SIMT)')
    # DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of
    # this same command (profiles/traffic.json: {"<workload>/<prec>": {"kernel", "bytes_per_launch", "source"}})
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            traffic = json.load(f).get('%s/%s' % (args.workload, args.prec), {}).get('bytes_per_launch')
    except Exception:  # noqa: BLE001
        traffic = None
    k_ms, k_cnt = prof[slot]
    k_ms_step = k_ms / args.steps
    achieved = fl / (k_ms_step * 1e-3) / 1e12 if k_ms_step > 0 else 0.0
    kernel_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1] > 0}
    line = {
        'metric': 'AEP energy+grad data-points/sec', 'value': value, 'unit': 'rows/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64' if pr == ops.F64 else 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload + ': ' + w['desc'], 'prec': args.prec,
                   'flops_per_row': flops_per_row(w), 'parallelism': 'dp%d (row-sharded, 1 packed all-reduce/step)' % world,
                   'l2': 'inputs and saved Kfu/T buffers (GBs) exceed the 126 MB L2; no flush needed'},
        'energy': energy,
        'clocks': clocks,
        'e2e': {'value': N / (ms_e2e * 1e-3), 'unit': 'rows/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'fp64_fma_pipe' if pr == ops.F64 else 'fp32_fma_pipe', 'kernel': kname,
                     'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': achieved / peak_tf if peak_tf > 0 else None, 'traffic': traffic,
                     'peak_source': 'gpb_fma_peak microbenchmark measured in this run (MEASURED_PEAKS.json has '
                                    'no FP64/FP32 FMA figure)',
                     'kernel_ms_per_step': k_ms_step, 'kernel_launches_per_step': k_cnt / max(args.steps, 1),
                     'algorithmic_flops_per_step': fl,
                     'whole_step_frac': flops_per_row(w) * N / world / (ms_step * 1e-3) / 1e12 / peak_tf},
        'kernel_ms_per_step': kernel_ms,
    }
    if world == 1 and not args.no_cpu:
        r = time_cpu(w, budget_s=10.0)
        line['cpu_baseline'] = {'value': r['value'], 'unit': 'rows/s', 'cores': r['cores'],
                                'threads_available': r['threads_available'], 'kind': r['kind'],
                                'sample': r['sample']}
    print(json.dumps(line))
    if world > 1:
        tdist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg3_sdgpr', choices=sorted(WORKLOADS))
    ap.add_argument('--prec', default='fp64', choices=['fp64', 'fp32'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    # Native libraries (NCCL's version banner, ...) write to fd 1 behind Python's back; the contract
    # is ONE JSON line on stdout, so fd 1 is pointed at stderr and Python keeps the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w')
    if args.impl == 'reference':
        run_reference(args, w)
    else:
        run_gpu(args, w)


if __name__ == '__main__':
    main()
