"""get_sorted_foreground at the inference shapes: 297 472 points per frame (64 x (2656 + 1328 + 664)), top 50 000."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import ops
N, K = 64 * (2656 + 1328 + 664), 50000
g = torch.Generator(device="cuda").manual_seed(0)
for B in (1, 8):
    score = torch.rand((B, N), device="cuda", generator=g)
    delta = torch.randn((B, N, 8), device="cuda", generator=g)
    pc = torch.randn((B, N, 3), device="cuda", generator=g)
    mask = (torch.rand((B, N), device="cuda", generator=g) > 0.3).float()
    for _ in range(3): ops.get_sorted_foreground(score, delta, pc, mask, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.get_sorted_foreground(score, delta, pc, mask, K)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"B": B, "points": N, "top": K, "ms": round(a.elapsed_time(b) / 20, 4)}), flush=True)
