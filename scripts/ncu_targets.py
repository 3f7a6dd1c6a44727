"""Targets for the round-2 ncu captures (run under `ncu --profile-from-start off`): only the region between
cudaProfilerStart / Stop is profiled.
    python scripts/ncu_targets.py step      # two eager cfg-5 train steps (B=2, fp16): launch list
    python scripts/ncu_targets.py kernels   # the dominant kernels of that step, one shape each, launched alone twice"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangedet_b200 import ops, synth, train  # noqa: E402
from rangedet_b200.model_params import make_params  # noqa: E402

H, W, B, DT = 64, 2656, 2, torch.float16
mode = sys.argv[1] if len(sys.argv) > 1 else "step"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = torch.Generator(device=dev).manual_seed(0)

if mode == "step":
    # eager launches (capture=False: same buffers, flat plumbing and kernels as the captured step): ncu reports LaunchFailed
    # when it replays kernel nodes of the captured graphs one by one
    step = train.GraphedTrainStep(make_params(seed=0, device=dev), B, H, W, lr=0.0125, device=dev, act_dtype=DT, capture=False)
    step.set_targets(synth.rpn_targets(B, seed=500))
    data = torch.randn((B, 8, H, W), device=dev, generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=700)).to(dev)
    for _ in range(2):
        step.train_step(data, coord)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(2):
        step.train_step(data, coord)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("launches per step", step.launches)
else:
    rnd = lambda *s: torch.randn(s, device=dev, generator=g)
    pad = lambda t: ops.to_nhwc_padded(t, dtype=DT)
    cases = []
    for ci, co, w in ((128, 128, 2656), (64, 64, 2656), (128, 128, 664)):
        x, wt = pad(rnd(B, ci, H, w)), ops.pack_conv_weight(rnd(co, ci, 3, 3) * 0.03, dtype=DT)
        y = torch.zeros((B, H + 2, w + 2, co), device=dev, dtype=DT)
        cases.append(lambda x=x, wt=wt, y=y: ops.conv2d_nhwc(x, wt, relu=False, out=y))            # dgrad form
        cases.append(lambda x=x, wt=wt, y=y: ops.conv2d_nhwc_stats(x, wt, out=y))                   # training forward + statistics
        dz = pad(rnd(B, co, H, w))
        cases.append(lambda dz=dz, x=x: ops.conv2d_wgrad(dz, x, 3, 1))
    z, dy = pad(rnd(B, 128, H, 2656)), pad(rnd(B, 128, H, 2656))
    coef = ops.bn_train_stats(z, torch.ones(128, device=dev), torch.zeros(128, device=dev))
    yb, dzb = torch.zeros_like(z), torch.zeros_like(z)
    cases.append(lambda: ops.bn_act_fwd(z, coef, relu=True, out=yb))
    cases.append(lambda: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dzb))
    dets = torch.from_numpy(synth.wnms_dets(100000, seed=0, clustered=True)).to(dev)
    cases.append(lambda: ops.wnms_4c_device(dets, 0.1, 0.5, False, 100))                          # cfg-3 weighted NMS, 100 k boxes
    for c in cases:   # warm (function attributes, workspaces)
        c()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for c in cases:
        c()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
