"""Data-parallel plumbing for the sharded hot path (one process per GPU, torch.distributed).

The path shards by frames: each rank owns a contiguous slice of the (shuffled) frame index, like
the reference's loader (utils/detection_input.py:49-54,117-122).  Forward/backward need no
collective (BatchNorm is local, config:56); the only exchange is the gradient all-reduce that
hvd.DistributedOptimizer performs per parameter (tools/train.py:364-368), done here as ONE flat
all-reduce (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous rank slice [lo, hi) of n_items; the first n_items % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_grads(grads):
    return torch.cat([g.reshape(-1) for g in grads])


def unflatten_like(flat, like):
    out, o = [], 0
    for g in like:
        n = g.numel()
        out.append(flat[o:o + n].reshape(g.shape))
        o += n
    return out


def allreduce_mean_(flat, world=None):
    """In-place average of a flat gradient buffer over all ranks (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized():
        w = dist.get_world_size() if world is None else world
        if w > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(w)
    return flat
