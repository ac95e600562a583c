"""Data-parallel path on CPU: world_size=2 over gloo, each rank running the emulated kernels on
its slice of the minibatch rows, ONE packed all-reduce of the statistics, replicated tail.
The 2-rank result must equal the golden (single-process reference) result."""
import copy
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, names, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, '..'))
    sys.path.insert(0, os.path.join(HERE, '..', 'oracle'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    import emu_util
    import golden_util as gu
    import model_cases as mc
    emu_util.attach()
    try:
        for name in names:
            mc.check_model(name, 'fp64', 1e-6)       # every rank checks the full result
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, 'FAIL %s: %r' % (name, e)))
    finally:
        torch.distributed.destroy_process_group()


@pytest.mark.parametrize('names', [
    ['aep_sgpr', 'aep_sgpr_minibatch', 'aep_sdgpr', 'vfe_sgpr', 'aep_sdgprh', 'aep_sdgprh_alpha_one'],
    ['aep_sgplvm', 'aep_sgplvm_minibatch', 'aep_sgpssm_lin', 'aep_sgpssm_lin_window', 'aep_sgpssm_gp',
     'vfe_sgplvm', 'vfe_sgpssm_lin', 'aep_sgplvm_mc', 'aep_sgplvm_mc_minibatch', 'vfe_sgplvm_mc', 'aep_sgpssm_lin_mc',
     'aep_sgpssm_gp_mc', 'aep_sgpssm_control_mc'],
])
def test_two_ranks_match_golden(names):
    import emu_util
    emu_util.attach()          # build the emulator once, before forking
    emu_util.detach()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, names, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 'ok' for r in res), res


def test_shard_covers_rows():
    from geepee_b200 import dist
    assert dist.shard(10) == (0, 10)


def _worker_rng(rank, world, port, q):
    """Ranks with DIFFERENT numpy RNG states must still shard the same minibatch / window / eps."""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, '..'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from geepee_b200 import dist
        np.random.seed(1000 + rank)                     # unsynchronised on purpose
        rows = dist.agree(np.random.choice(50, 7, replace=False))
        start = int(dist.agree(np.random.randint(0, 40)))
        eps = dist.agree(np.random.randn(2, 5, 3))
        q.put((rank, rows.tolist(), start, float(eps.sum())))
    finally:
        torch.distributed.destroy_process_group()


def test_random_draws_agree_across_ranks():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_rng, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1:] == res[1][1:], res
    ref = np.random.RandomState(1000).choice(50, 7, replace=False).tolist()     # rank 0's own draw
    assert res[0][1] == ref
