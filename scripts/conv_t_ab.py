"""A/B of the two conv orientations on the step's 128-channel 3x3 shapes (B=2, fp16): CUDA events, 20 launches each."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import _lib, ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
for (ci, w, co) in [(128, 2656, 128), (128, 664, 128), (64, 2656, 64), (64, 1328, 64), (128, 1328, 64)]:
    B, H = 2, 64
    x = ops.to_nhwc_padded(torch.randn((B, ci, H, w), device=dev, generator=g), dtype=DT)
    wt = ops.pack_conv_weight(torch.randn((co, ci, 3, 3), device=dev, generator=g) * 0.03, dtype=DT)
    y = torch.zeros((B, H + 2, w + 2, co), device=dev, dtype=DT)
    res = {}
    for on in (True, False):
        _lib.set_conv_t(on)
        for fn_name, fn in (("plain", lambda: ops.conv2d_nhwc(x, wt, relu=False, out=y)), ("stats", lambda: ops.conv2d_nhwc_stats(x, wt, out=y))):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): fn()
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 20 * 1e3
            res["%s_%s" % ("T" if on else "P", fn_name)] = {"us": round(us, 1), "TFLOPs": round(2.0 * B * H * w * ci * co * 9 / us / 1e6, 1)}
    _lib.set_conv_t(True)
    print(json.dumps({"Cin": ci, "Cout": co, "W": w, **res}), flush=True)
