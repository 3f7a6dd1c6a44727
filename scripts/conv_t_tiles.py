"""Pixel-tile width of the transposed-orientation conv (csrc/conv_t.cu) per layer width of the step (B=2, fp16, 128 -> 128, 3x3):
fixed 160 / 192 / 224 / 256 against the per-shape choice (`auto`).  CUDA events, 30 back-to-back launches each (the
input, 11-90 MB, stays L2-resident for the narrow layers exactly as it does inside the step)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import _lib, ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
B, H, ci, co = 2, 64, 128, 128
for w in (2656, 1328, 664, 332, 166):
    x = ops.to_nhwc_padded(torch.randn((B, ci, H, w), device=dev, generator=g), dtype=DT)
    wt = ops.pack_conv_weight(torch.randn((co, ci, 3, 3), device=dev, generator=g) * 0.03, dtype=DT)
    y = torch.zeros((B, H + 2, w + 2, co), device=dev, dtype=DT)
    res = {}
    for rep in range(2):
        for tn in (1, 160, 192, 224, 256):
            _lib.set_conv_t(tn)
            fn = lambda: ops.conv2d_nhwc_stats(x, wt, out=y)
            for _ in range(3): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30): fn()
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / 30 * 1e3
            k = "auto" if tn == 1 else str(tn)
            res[k] = round(min(us, res.get(k, 1e9)), 2)
    _lib.set_conv_t(1)
    print(json.dumps({"W": w, "us": res, "TFLOPs_auto": round(2.0 * B * H * w * ci * co * 9 / res["auto"] / 1e6, 1)}), flush=True)
