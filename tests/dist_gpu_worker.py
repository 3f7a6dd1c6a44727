"""Worker of tests/test_gpu_dist.py: launched as `torchrun --nproc-per-node N tests/dist_gpu_worker.py`, one rank
per GPU over NCCL.  Checks the three exchanges of SURVEY 8(e) on the real kernels and prints ONE JSON line (rank 0):

  start-up broadcast  (tools/train.py:219-229)            ranks initialised with different seeds end up bit-identical
  gradient all-reduce (tools/train.py:364-368)            flat buffer after the NCCL sum == g_rank0 + g_rank1 + ... bit for bit,
                                                          where each g_r is ALSO recomputed by rank 0 alone from rank r's frames
  update                                                  parameters bit-identical on all ranks after the SGD step
  epoch-end aux average (utils/detection_module.py:1164-1170)   moving statistics == mean over ranks, identical everywhere
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from rangedet_b200 import synth, train
    from rangedet_b200.model_params import make_params
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    B, H, W = 1, 8, 256
    res = {"world": world}

    def gather(t):
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    def allreduce(flat):       # asynchronous: the step overlaps the head bucket with the backbone's backward
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)

    def local_backward():
        """loss + backward of the frames already in the static buffers, no exchange"""
        if step.split_bwd:
            for g in step.g_bwd_parts:
                g.replay()
        else:
            step.g_bwd.replay()

    P = make_params(seed=10 + rank, device=dev)               # every rank initialises differently
    for k in P:                                               # ... including the moving statistics
        if k.endswith(("_moving_mean", "_moving_var")):
            P[k].add_(0.01 * rank)
    step = train.GraphedTrainStep(P, B, H, W, lr=0.05, device=dev, allreduce=allreduce, world_size=world)
    before = gather(step.flatP)
    step.broadcast_parameters(src=0)
    after = gather(step.flatP)
    aux_names = sorted(k for k in P if k.endswith(("_moving_mean", "_moving_var")))
    aux = gather(torch.cat([P[k].reshape(-1) for k in aux_names]))
    res["init_differs"] = not torch.equal(before[0], before[-1])
    res["bcast_args_equal"] = all(torch.equal(after[0], a) for a in after) and torch.equal(after[0], before[0])
    res["bcast_aux_equal"] = all(torch.equal(aux[0], a) for a in aux) and bool(float(aux[0].max()) <= 1.0)   # rank 0's values

    def frames(r):
        T = synth.rpn_targets(B, seed=500 + r, n_vehicles=4, h=H, w=W - 6, w_pad=W)
        g = torch.Generator(device=dev).manual_seed(600 + r)
        data = torch.randn((B, 8, H, W), device=dev, generator=g)
        xyz = torch.from_numpy(T["pc_vehicle_frame_s1"]).to(dev).reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous() / 25.0
        return T, data, xyz

    def local_grad(r):
        """forward + loss + backward of rank r's frames on THIS rank, no exchange, no update."""
        T, data, xyz = frames(r)
        step.set_targets(T)
        step.forward(data, xyz)
        local_backward()
        torch.cuda.synchronize()
        return step.flat.clone()

    aux0 = {k: P[k].clone() for k in aux_names}
    g_mine = local_grad(rank)
    g_all = gather(g_mine)
    want_sum = g_all[0].clone()
    for g in g_all[1:]:
        want_sum += g
    if rank == 0:   # the other ranks' gradients recomputed single-handed (same parameters, their frames)
        res["single_rank_recompute_bitexact"] = all(torch.equal(local_grad(r), g_all[r]) for r in range(1, world))
    for k in aux_names:   # the probing passes above must not count as training steps for the moving statistics
        P[k].copy_(aux0[k])
    dist.barrier()
    # the real step: forward, loss, backward, NCCL all-reduce, update
    T, data, xyz = frames(rank)
    step.set_targets(T)
    step.forward(data, xyz)
    res["split_backward"] = bool(step.split_bwd)
    hyper = step.hyper.clone()
    step.hyper[0] = 0.0                      # lr 0: the exchange as train_step runs it (two overlapped buckets), weights untouched
    step.backward_update()
    torch.cuda.synchronize()
    res["allreduce_bitexact_sum"] = bool(torch.equal(step.flat, want_sum))
    res["grad_nonzero"] = bool(float(want_sum.abs().max()) > 0)
    step.hyper.copy_(hyper)
    step.flat_m.zero_()
    step.g_upd.replay()                      # the real update on the exchanged gradients
    torch.cuda.synchronize()
    upd = gather(step.flatP)
    res["params_equal_after_update"] = all(torch.equal(upd[0], u) for u in upd)
    res["params_moved"] = not torch.equal(upd[0], after[0])
    # epoch end: moving statistics drifted apart (different frames per rank), then averaged
    cat_aux = lambda: torch.cat([P[k].reshape(-1) for k in aux_names])
    drift = gather(cat_aux())
    res["aux_drifted"] = not torch.equal(drift[0], drift[-1])
    step.average_aux()
    avg = gather(cat_aux())
    mean = sum(drift) / world
    res["aux_equal_after_average"] = all(torch.equal(avg[0], a) for a in avg)
    res["aux_average_err"] = float((avg[0] - mean).abs().max())
    # a second full train_step keeps the ranks in lock-step
    step.train_step(data, xyz)
    torch.cuda.synchronize()
    upd2 = gather(step.flatP)
    res["lockstep_second_step"] = all(torch.equal(upd2[0], u) for u in upd2) and bool(torch.isfinite(upd2[0]).all())
    if rank == 0:
        print("DIST_PARITY " + json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
