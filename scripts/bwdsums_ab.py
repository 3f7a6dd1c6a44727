"""Fused BatchNorm-backward sums: data-gradient conv (+ sums) and BN backward (apply only) against the separate passes,
per layer width of the head towers / backbone (B=2, fp16, 128 channels).  CUDA events, 20 launches of each pair."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
B, H, C = 2, 64, 128
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n * 1e3, 1)
for w in (2656, 1328, 664, 332, 166):
    dzu = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
    z = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
    wt = ops.pack_conv_weight(torch.randn((C, C, 3, 3), device=dev, generator=g) * 0.03, dtype=DT)
    coef = ops.bn_train_stats(z, torch.ones(C, device=dev), torch.zeros(C, device=dev))
    dy, dz = torch.zeros_like(z), torch.zeros_like(z)
    ws = torch.empty(1184 * 2 * C, device=dev)
    r = {"W": w}
    r["conv_us"] = timed(lambda: ops.conv2d_nhwc(dzu, wt, relu=False, out=dy))
    r["conv_sums_us"] = timed(lambda: ops.conv2d_nhwc_bwdstats(dzu, wt, z, coef, 2, out=dy, ws=ws))
    r["bn_bwd_us"] = timed(lambda: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz))
    _, s, n = ops.conv2d_nhwc_bwdstats(dzu, wt, z, coef, 2, out=dy, ws=ws)
    r["bn_bwd_apply_us"] = timed(lambda: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz, sums=(s, n)))
    r["pair_separate_us"] = timed(lambda: (ops.conv2d_nhwc(dzu, wt, relu=False, out=dy), ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz)))
    r["pair_fused_us"] = timed(lambda: (lambda t: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz, sums=(t[1], t[2])))(ops.conv2d_nhwc_bwdstats(dzu, wt, z, coef, 2, out=dy, ws=ws)))
    print(json.dumps(r), flush=True)
