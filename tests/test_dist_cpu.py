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


def _train_worker(rank, world, port, ret):
    """Whole-model data-parallel step on CPU: per-rank frames + LOCAL BatchNorm statistics (config:56), ONE flat
    SUM all-reduce of every parameter gradient, MXNet SGD with rescale_grad / world -- the exchange contract of
    rangedet_b200.train.GraphedTrainStep (tools/train.py:306-319, 359-368), with the torch restatement standing in
    for the kernels."""
    from oracle import dla_ref, dla_train_ref
    from rangedet_b200 import train
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    H, W = 8, 32
    P = dla_ref.make_params(seed=0, device="cpu")
    g = torch.Generator().manual_seed(5)
    data = torch.randn(world, 8, H, W, generator=g)
    coord = torch.from_numpy(synth.range_image_coords(world, seed=0, h=H, w=W - 2, w_pad=W))
    d_cls = [torch.randn(world, 1, H, W // s, generator=g) for s in (1, 2, 4)]
    d_reg = [torch.randn(world, 8, H, W // s, generator=g) for s in (1, 2, 4)]
    names = sorted(k for k in P if not k.endswith(("_moving_mean", "_moving_var")))

    def local_grads(r):
        tr = dla_train_ref.TrainRef(P, bf16=False)
        _, _, grads = tr.forward_backward(data[r:r + 1], coord[r:r + 1], [d[r:r + 1] for d in d_cls], [d[r:r + 1] for d in d_reg])
        return torch.cat([grads[k].reshape(-1) for k in names])

    flat = local_grads(rank)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                      # what GraphedTrainStep.allreduce must do
    Pw = {k: P[k].clone() for k in names}
    grads = {k: v for k, v in zip(names, rd_dist.unflatten_like(flat, [P[k] for k in names]))}
    train.sgd_momentum_step(Pw, grads, {}, lr=0.05, momentum=0.9, wd=1e-5, clip_gradient=35.0, rescale_grad=1.0 / 128 / world)
    after = torch.cat([Pw[k].reshape(-1) for k in names])
    both = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(both, after)
    if rank == 0:
        ret["ranks_agree"] = bool(torch.equal(both[0], both[1]))
        mean = sum(local_grads(r) for r in range(world)) / world     # single-process emulation of the same update
        Ps = {k: P[k].clone() for k in names}
        gs = {k: v for k, v in zip(names, rd_dist.unflatten_like(mean, [P[k] for k in names]))}
        train.sgd_momentum_step(Ps, gs, {}, lr=0.05, momentum=0.9, wd=1e-5, clip_gradient=35.0, rescale_grad=1.0 / 128)
        want = torch.cat([Ps[k].reshape(-1) for k in names])
        ret["err"] = float((after - want).abs().max())
        ret["moved"] = float((after - torch.cat([P[k].reshape(-1) for k in names])).abs().max())
        ret["n"] = int(after.numel())
    dist.destroy_process_group()


def test_gloo_world2_whole_model_exchange_and_update():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_train_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["ranks_agree"] and ret["n"] > 9_000_000
    assert ret["err"] < 1e-6 and ret["moved"] > 1e-4


def _sync_worker(rank, world, port, ret):
    """Start-up broadcast (tools/train.py:219-229) and epoch-end aux average (utils/detection_module.py:1164-1170)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)          # every rank initialises DIFFERENTLY
    mk = lambda *s: torch.randn(*s, generator=g)
    P = {"a_weight": mk(4, 3, 3, 3), "a_bn_gamma": mk(4), "a_bn_beta": mk(4), "a_bn_moving_mean": mk(4), "a_bn_moving_var": mk(4).abs(),
         "b_weight": mk(2, 4, 1, 1), "b_bias": mk(2)}
    mine = {k: v.clone() for k, v in P.items()}
    # (1) generic path: every tensor staged through one flat buffer
    rd_dist.broadcast_params_(P, src=0)
    gathered = {}
    for k in sorted(P):
        both = [torch.zeros_like(P[k]) for _ in range(world)]
        dist.all_gather(both, P[k])
        gathered[k] = both
    # (2) flat-master path: trainable parameters are views of one buffer (GraphedTrainStep layout)
    names = sorted(k for k in mine if not k.endswith(("_moving_mean", "_moving_var")))
    flat = torch.cat([mine[k].reshape(-1) for k in names])
    Q, o = {}, 0
    for k in names:
        n = mine[k].numel()
        Q[k] = flat[o:o + n].view(mine[k].shape)
        o += n
    for k in mine:
        if k not in Q:
            Q[k] = mine[k].clone()
    rd_dist.broadcast_params_(Q, src=0, flat=flat)
    same_as_generic = all(torch.equal(Q[k], P[k]) for k in P)
    # (3) aux average: moving statistics drift apart per rank, then are averaged; trainable parameters untouched
    P["a_bn_moving_mean"] += rank + 1.0
    P["a_bn_moving_var"] *= rank + 2.0
    before = {k: v.clone() for k, v in P.items()}
    both_mm = [torch.zeros(4) for _ in range(world)]
    dist.all_gather(both_mm, P["a_bn_moving_mean"])
    rd_dist.average_aux_(P)
    if rank == 0:
        ret["bcast_equal"] = all(torch.equal(v[0], v[1]) for v in gathered.values())
        ret["bcast_is_rank0"] = all(torch.equal(gathered[k][0], mine[k]) for k in mine)
        ret["flat_path_same"] = same_as_generic
        ret["aux_avg_err"] = float((P["a_bn_moving_mean"] - (both_mm[0] + both_mm[1]) / 2).abs().max())
        ret["args_untouched"] = all(torch.equal(P[k], before[k]) for k in names)
    else:
        ret["rank1_changed"] = not torch.equal(gathered["a_weight"][1], mine["a_weight"])
    dist.destroy_process_group()


def test_gloo_world2_startup_broadcast_and_aux_average():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sync_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret["bcast_equal"] and ret["bcast_is_rank0"] and ret["rank1_changed"] and ret["flat_path_same"]
    assert ret["aux_avg_err"] < 1e-6 and ret["args_untouched"]
