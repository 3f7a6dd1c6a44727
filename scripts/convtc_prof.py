import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
B, H = 2, 64
for (ci, co, w, k, s) in ((64, 64, 2656, 3, 1), (64, 128, 2656, 3, 2), (576, 64, 2656, 1, 1), (64, 128, 2656, 1, 1)):
    x = ops.to_nhwc_padded(torch.randn((B, ci, H, w), device=dev, generator=g), dtype=DT)
    wt = ops.pack_conv_weight(torch.randn((co, ci, k, k), device=dev, generator=g) * 0.03, dtype=DT)
    print("== %dx%d %d->%d @%d s%d" % (k, k, ci, co, w, s), file=sys.stderr, flush=True)
    for _ in range(2):
        ops.conv2d_nhwc(x, wt, relu=False, stride_w=s)
    torch.cuda.synchronize()
