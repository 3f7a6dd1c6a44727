import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import ops
dev = "cuda"
res = {}
for (n, cin, cout, h, w) in [(4, 64, 64, 64, 2656), (4, 128, 128, 64, 2656)]:
    x = ops.to_nhwc_padded(torch.randn(n, cin, h, w, device=dev))
    wt = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    y = torch.zeros((n, h + 2, w + 2, cout), device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.conv2d_nhwc(x, wt, None, None, relu=True, out=y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.conv2d_nhwc(x, wt, None, None, relu=True, out=y)
    b.record(); torch.cuda.synchronize()
    res["%d_%d" % (cin, cout)] = a.elapsed_time(b) / 20
print(json.dumps(res))
