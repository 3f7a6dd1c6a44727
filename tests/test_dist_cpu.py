"""world_size-2 gloo test (CPU) of the N>1 host logic: frame sharding + flat gradient all-reduce.
Local gradients come from the CPU oracle (the GPU kernels are covered by -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rangedet_b200 import dist as rd_dist
from rangedet_b200 import synth


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [rd_dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rd_dist.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    from oracle import meta_kernel_ref
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    B, C, H, W = 4, 64, 4, 36
    data = torch.from_numpy(synth.feature_map(B, C, seed=1, h=H, w=W - 2, w_pad=W))
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0, h=H, w=W - 2, w_pad=W))
    ps = [torch.from_numpy(p) for p in synth.meta_mlp_params(seed=2)]
    go = torch.from_numpy(np.random.default_rng(3).standard_normal((B, 9 * C, H, W)).astype(np.float32))
    lo, hi = rd_dist.shard_range(B, rank, world)
    res = meta_kernel_ref.meta_baseline_bias_fwd_bwd(data[lo:hi], coord[lo:hi], *ps, go[lo:hi])
    flat = rd_dist.flatten_grads(res[2:])
    rd_dist.allreduce_mean_(flat)
    if rank == 0:
        full = meta_kernel_ref.meta_baseline_bias_fwd_bwd(data, coord, *ps, go)
        want = rd_dist.flatten_grads(full[2:]) / world
        ret["err"] = float((flat - want).abs().max() / want.abs().max())
        ret["shapes"] = [tuple(g.shape) for g in rd_dist.unflatten_like(flat, res[2:])]
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_matches_global_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["err"] < 1e-5
    assert ret["shapes"] == [(32, 3), (32,), (64, 32), (64,)]
