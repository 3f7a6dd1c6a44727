import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import ops
box_w, c0, c1, c2 = [int(x) for x in sys.argv[1:5]]
C, H, W = 128, 5, 264
src = torch.randn(C, H, W, device="cuda")
try:
    tile, back = ops.tma_probe(src, box_w, c0, c1, c2)
    torch.cuda.synchronize()
    want = torch.zeros(64, box_w, device="cuda"); wb = torch.zeros_like(src)
    if 0 <= c1 < H:
        lo, hi = max(c0, 0), min(c0 + box_w, W)
        if hi > lo:
            want[:, lo - c0:hi - c0] = src[c2:c2 + 64, c1, lo:hi]; wb[c2:c2 + 64, c1, lo:hi] = src[c2:c2 + 64, c1, lo:hi]
    print("CASE", sys.argv[1:5], "load_ok", bool(torch.equal(tile, want)), "store_ok", bool(torch.equal(back, wb)))
except Exception as e:
    print("CASE", sys.argv[1:5], "ERROR", str(e)[:80].replace("\n", " "))
