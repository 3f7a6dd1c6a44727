"""Same-box A/B of the round-2 step changes: fused conv statistics and programmatic dependent launch, B = 2 and 4.
    python scripts/ab_step.py  ->  one JSON line per configuration (ms per captured step, 20 replays after 5 warm-up)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangedet_b200 import _lib, synth, train  # noqa: E402
from rangedet_b200.model_params import make_params  # noqa: E402

H, W = 64, 2656
dev = torch.device("cuda", 0)
res = []
for B in (2, 4):
    T = synth.rpn_targets(B, seed=500)
    g = torch.Generator(device=dev).manual_seed(600)
    data = torch.randn((B, 8, H, W), device=dev, generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=700)).to(dev)
    for rep in range(2):
        for fuse, pdl in ((1, 1), (0, 1), (1, 0), (0, 0)):
            _lib.set_pdl(bool(pdl))
            step = train.GraphedTrainStep(make_params(seed=0, device=dev), B, H, W, lr=0.0125, device=dev, act_dtype=torch.float16,
                                          fuse_stats=bool(fuse))
            step.set_targets(T)
            for _ in range(5):
                step.train_step(data, coord)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                step.train_step(data, coord)
            b.record()
            torch.cuda.synchronize()
            r = {"B": B, "rep": rep, "fuse_stats": fuse, "pdl": pdl, "ms_per_step": a.elapsed_time(b) / 20, "launches": step.launches}
            print(json.dumps(r), flush=True)
            del step
            torch.cuda.empty_cache()
_lib.set_pdl(True)
