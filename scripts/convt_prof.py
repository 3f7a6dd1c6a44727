import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
B, H, C, w = 2, 64, 128, 2656
x = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
r = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
wt = ops.pack_conv_weight(torch.randn((C, C, 3, 3), device=dev, generator=g) * 0.03, dtype=DT)
coef = ops.bn_train_stats(r, torch.ones(C, device=dev), torch.zeros(C, device=dev))
y = torch.zeros_like(r)
ws = torch.empty(1184 * 2 * C, device=dev)
for name, fn in (("plain", lambda: ops.conv2d_nhwc(x, wt, relu=False, out=y)),
                 ("residual", lambda: ops.conv2d_nhwc(x, wt, relu=False, residual_pad=r, out=y)),
                 ("bwdsums", lambda: ops.conv2d_nhwc_bwdstats(x, wt, r, coef, 2, out=y, ws=ws))):
    print("==", name, file=sys.stderr, flush=True)
    for _ in range(2): fn()
    torch.cuda.synchronize()
