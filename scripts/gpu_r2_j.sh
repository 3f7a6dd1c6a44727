#!/bin/bash
mkdir -p gpurun_out
RD_CONVT_PROF=1 timeout 300 python - <<'PY' 2>&1 | tail -12
import sys, torch
sys.path.insert(0, ".")
from rangedet_b200 import ops
DT = torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
for (ci, w) in [(128, 2656), (128, 664), (64, 1328), (256, 664)]:
    x = ops.to_nhwc_padded(torch.randn((2, ci, 64, w), device="cuda", generator=g), dtype=DT)
    wt = ops.pack_conv_weight(torch.randn((128, ci, 3, 3), device="cuda", generator=g) * 0.03, dtype=DT)
    y = torch.zeros((2, 66, w + 2, 128), device="cuda", dtype=DT)
    for _ in range(3):
        ops.conv2d_nhwc(x, wt, relu=False, out=y)
    torch.cuda.synchronize()
PY
