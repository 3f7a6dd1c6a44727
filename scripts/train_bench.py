"""Whole-model training step (backbone + Meta-Kernel unit + head, training-mode BN, forward + backward +
SGD update) on one B200: frames/s at the shipped per-GPU batch (B=2, config:32) and at B=4.
Synthetic data of the reference's shapes, random-init weights (oracle-free: parameters are generated
here with the reference's names and shapes).  Timing: CUDA events on the launching stream, after warm-up.

    python scripts/train_bench.py [--batch 2 4] [--steps 10] [--warmup 3] [--no-meta]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from rangedet_b200 import _lib, synth, train  # noqa: E402
from rangedet_b200.model_params import make_params  # noqa: E402

FLOP_FWD_PER_FRAME = 1.114e12  # backbone 480 + head 634 GFLOP forward per 64x2656 frame


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[2, 4])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-meta", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    ap.add_argument("--no-overlap", action="store_true", help="weight gradients on the main stream (no second stream)")
    ap.add_argument("--no-loss", action="store_true", help="linear loss (random head-output gradients) instead of the RPN loss")
    a = ap.parse_args()
    H, W = 64, 2656
    out = []
    for B in a.batch:
        P = make_params(seed=0, device="cuda")
        tg = train.TrainGraph(P, use_meta=not a.no_meta)
        mom = {}
        g = torch.Generator(device="cuda").manual_seed(1)
        data = torch.randn((B, 8, H, W), device="cuda", generator=g)
        coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).cuda()
        Ws = [W, W // 2, W // 4]
        d_cls = [torch.randn((B, 1, H, w), device="cuda", generator=g) * 1e-3 for w in Ws]
        d_reg = [torch.randn((B, 8, H, w), device="cuda", generator=g) * 1e-3 for w in Ws]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        t_f = t_b = t_u = 0.0
        l0 = _lib.launch_count()
        wall0 = None
        step = None if a.eager else train.GraphedTrainStep(P, B, H, W, lr=1e-4, clip_gradient=35.0, use_meta=not a.no_meta,
                                                           with_loss=not a.no_loss, overlap_wgrad=not a.no_overlap)
        if step is not None and not a.no_loss:
            step.set_targets(synth.rpn_targets(B, seed=5))
        for it in range(a.warmup + a.steps):
            if step is not None:
                if it == a.warmup:
                    torch.cuda.synchronize()
                    l0 = _lib.launch_count()
                    wall0 = time.perf_counter()
                ev[0].record()
                step.forward(data, coord)
                ev[1].record()
                step.backward_update(d_cls, d_reg) if a.no_loss else step.backward_update()
                ev[2].record()
                ev[3].record()
                if it >= a.warmup:
                    torch.cuda.synchronize()
                    t_f += ev[0].elapsed_time(ev[1])
                    t_b += ev[1].elapsed_time(ev[2])
                continue
            if it == a.warmup:
                torch.cuda.synchronize()
                l0 = _lib.launch_count()
                wall0 = time.perf_counter()
            ev[0].record()
            tg.forward(data, coord)
            ev[1].record()
            grads = tg.backward(d_cls, d_reg)
            ev[2].record()
            train.sgd_momentum_step(P, grads, mom, lr=1e-4, clip_gradient=35.0)
            tg.refresh()
            ev[3].record()
            if it >= a.warmup:
                torch.cuda.synchronize()
                t_f += ev[0].elapsed_time(ev[1])
                t_b += ev[1].elapsed_time(ev[2])
                t_u += ev[2].elapsed_time(ev[3])
        torch.cuda.synchronize()
        wall = (time.perf_counter() - wall0) / a.steps * 1e3
        loss_ms = None
        if step is not None and not a.no_loss:   # the loss head alone (3 levels, 6 launches), same buffers
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            train.rpn_loss_levels(step.out[0], step.out[1], step.targets, out=step.loss_out)
            e0.record()
            for _ in range(5):
                train.rpn_loss_levels(step.out[0], step.out[1], step.targets, out=step.loss_out)
            e1.record()
            torch.cuda.synchronize()
            loss_ms = e0.elapsed_time(e1) / 5
        n = a.steps
        step_ms = (t_f + t_b + t_u) / n
        finite = all(bool(torch.isfinite(v).all()) for v in P.values())
        out.append({"batch": B, "fwd_ms": t_f / n, "bwd_ms": t_b / n, "update_ms": t_u / n, "loss_ms": loss_ms, "step_ms": step_ms,
                    "wall_ms_per_step": wall, "frames_per_s": B / step_ms * 1e3,
                    "algorithmic_TFLOPs": 3 * FLOP_FWD_PER_FRAME * B / step_ms / 1e9,
                    "launches_per_step": (_lib.launch_count() - l0) / n,
                    "mem_GB": torch.cuda.max_memory_allocated() / 1e9, "params_finite": finite,
                    "meta_unit": not a.no_meta, "overlap_wgrad": not a.no_overlap, "loss": "linear" if (a.no_loss or a.eager) else "RPN loss (IoU target + VFL + smooth-L1)",
                    "mode": "eager launches" if a.eager else "CUDA graph replay (forward | loss + backward | update)"})
        del tg, P, mom, step
        torch.cuda.empty_cache()
    print(json.dumps({"workload": "train step fwd+bwd+SGD, DLA backbone + Meta-Kernel unit + RPN head, 64x2656, bf16 operands / "
                                  "fp32 accumulate, training-mode BN, synthetic roidb",
                      "results": out}))


if __name__ == "__main__":
    main()
