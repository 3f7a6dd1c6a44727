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


def _is_aux(name):
    return name.endswith(("_moving_mean", "_moving_var"))


def _flat_roundtrip(tensors, fn):
    """Apply the collective `fn` to ONE flat buffer holding `tensors` (sorted-key order is the caller's job) and
    copy the result back: one launch-latency-sized message instead of one per parameter."""
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    fn(flat)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view(t.shape))
        o += n


def broadcast_params_(params, src=0, flat=None):
    """Start-up synchronisation (tools/train.py:219-229: hvd.broadcast_parameters(arg_params, root_rank=0) and the
    same for aux_params): every rank ends up with rank `src`'s trainable parameters AND BatchNorm moving statistics,
    whatever it initialised locally.  `flat`: when the trainable parameters are views of one flat buffer
    (train.GraphedTrainStep.flatP) that buffer is broadcast in place and only the aux states are staged.
    No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return params
    bc = lambda t: dist.broadcast(t, src=src)
    names = sorted(params)
    if flat is not None:
        bc(flat)
        _flat_roundtrip([params[k] for k in names if _is_aux(k)], bc)
    else:
        _flat_roundtrip([params[k] for k in names], bc)
    return params


def average_aux_(params):
    """Per-epoch average of the BatchNorm moving statistics over ranks (utils/detection_module.py:1144,1164-1170:
    sync_params -> hvd.allreduce_(v, average=True) over aux_params; BatchNorm itself stays per-GPU, config:56).
    One flat all-reduce.  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return params
    w = dist.get_world_size()

    def avg(flat):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(w)

    _flat_roundtrip([params[k] for k in sorted(params) if _is_aux(k)], avg)
    return params
