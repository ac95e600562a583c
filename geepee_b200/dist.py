"""Data parallelism over minibatch rows.

One process per GPU.  Every rank holds the (small) parameters and the training arrays;
each evaluates the per-row kernels on its contiguous slice of the minibatch and the ranks
exchange ONE packed fp64 buffer of additive sufficient statistics per objective call
(`all_reduce(sum)`: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).  The
O(Dout M^3) tail then runs redundantly and identically on every rank, so no broadcast of the
result is needed.  With world size 1 everything here is a no-op.
"""
import torch


def world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def shard(n):
    """Contiguous slice [lo, hi) of `n` rows owned by this rank."""
    rank, ws = world()
    return (rank * n) // ws, ((rank + 1) * n) // ws


def agree(arr):
    """Host-side random draws (minibatch rows, SSM window, Monte-Carlo eps) must be THE SAME array on every
    rank: each rank draws from its own numpy RNG in the reference's order, then rank 0's draw is broadcast
    so that unsynchronised RNG states cannot make the ranks shard different minibatches.  No-op with one rank."""
    rank, ws = world()
    if ws == 1:
        return arr
    box = [arr]
    torch.distributed.broadcast_object_list(box, src=0)
    return box[0]


def allreduce_packed(tensors):
    """Sum a list of same-dtype device tensors across ranks with one collective.
    Returns new tensors (views into the packed buffer)."""
    rank, ws = world()
    if ws == 1:
        return tensors
    from . import tail
    flat = tail.gather([t.reshape(-1) for t in tensors])          # one library launch, no torch.cat
    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
    out, off = [], 0
    for t in tensors:
        k = t.numel()
        out.append(flat[off:off + k].reshape(t.shape))
        off += k
    return out


def gather_rows(full, n):
    """`full[n, ...]` holds valid rows only in this rank's contiguous shard [lo, hi) of `n` rows (the
    latent-variable gradients stay sharded with the rows): make every rank hold all rows with one broadcast per
    shard, N rows of traffic per rank instead of the all-reduce of a full-size array that is mostly zeros."""
    rank, ws = world()
    if ws == 1:
        return full
    for r in range(ws):
        lo, hi = (r * n) // ws, ((r + 1) * n) // ws
        if hi > lo:
            torch.distributed.broadcast(full[lo:hi], src=r)
    return full


def allreduce_dict(d):
    keys = sorted(d.keys())
    vals = allreduce_packed([d[k] for k in keys])
    return dict(zip(keys, vals))
