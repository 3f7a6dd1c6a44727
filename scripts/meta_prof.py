"""Per-role cycle profile of the Meta-Kernel impl-3 kernels (RD_MK_PROF=1 must be set): eager launches at the
BASELINE configs[1] shape; the library prints one line per launch to stderr."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import ops, synth
B, C, H, W = 4, 64, 64, 2656
g = torch.Generator(device="cuda").manual_seed(0)
data = torch.randn(B, C, H, W, device="cuda", generator=g)
coord = torch.randn(B, 3, H, W, device="cuda", generator=g) * 10
w0 = torch.randn(32, 3, device="cuda", generator=g) * 0.3
b0 = torch.randn(32, device="cuda", generator=g) * 0.1
w1 = torch.randn(C, 32, device="cuda", generator=g) * 0.2
b1 = torch.randn(C, device="cuda", generator=g) * 0.1
for _ in range(3):
    out = ops.meta_kernel_forward(data, coord, w0, b0, w1, b1, impl=3)
go = torch.randn_like(out)
for _ in range(3):
    ops.meta_kernel_backward(go, data, coord, w0, b0, w1, b1, impl=3)
torch.cuda.synchronize()
