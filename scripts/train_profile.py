"""In-situ kernel breakdown of the captured training step (CUPTI via torch.profiler on CUDA-graph replays; unlike
an ncu launch list the kernels run back to back with a warm L2, as in the timed step).
    python scripts/train_profile.py [batch] > profiles/..._breakdown.txt"""
import os
import re
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangedet_b200 import synth, train  # noqa: E402
from rangedet_b200.model_params import make_params  # noqa: E402


def short(name):
    m = re.search(r"([A-Za-z_0-9]+::)?([A-Za-z_0-9]+)\s*(<[^(]*)?\(", name)
    s = (m.group(1) or "") + m.group(2) if m else name[:50]
    if "conv_kernel" in name or "wgrad" in name:
        return s
    return s


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H, W = 64, 2656
    P = make_params(seed=0, device="cuda")
    step = train.GraphedTrainStep(P, B, H, W, lr=1e-4)
    step.set_targets(synth.rpn_targets(B, seed=5))
    g = torch.Generator(device="cuda").manual_seed(1)
    data = torch.randn((B, 8, H, W), device="cuda", generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).cuda()
    for _ in range(3):
        step.train_step(data, coord)
    torch.cuda.synchronize()
    nrep = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(nrep):
            step.train_step(data, coord)
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time_total > 0:
            k = short(ev.name)
            agg[k][0] += 1
            agg[k][1] += ev.device_time_total
    tot = sum(v[1] for v in agg.values())
    print("batch %d: %d kernel/memcpy records over %d steps, sum of durations %.3f ms per step" % (B, sum(v[0] for v in agg.values()), nrep, tot / nrep / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s n/step=%6.1f  %9.1f us/step  %5.1f %%  avg %8.1f us" % (k[:48], v[0] / nrep, v[1] / nrep, 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main()
