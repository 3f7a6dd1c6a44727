import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rangedet_b200 import ops, synth
from oracle import meta_kernel_ref
B, C, H, W = 1, 64, 5, 256
coord = synth.range_image_coords(B, seed=0, h=H, w=W - 4, w_pad=W)
data = synth.feature_map(B, C, seed=1, h=H, w=W - 4, w_pad=W)
w0, b0, w1, b1 = synth.meta_mlp_params(seed=2)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if mode == "fwd":
    out = ops.meta_kernel_forward(cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), impl=3)
    torch.cuda.synchronize()
    want = meta_kernel_ref.meta_baseline_bias(*[torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1)])
    print("fwd rel err", float((out.cpu() - want).abs().max() / want.abs().max()))
else:
    go = np.random.default_rng(3).standard_normal((B, 9 * C, H, W)).astype(np.float32)
    g = ops.meta_kernel_backward(cu(go), cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), impl=3)
    torch.cuda.synchronize()
    want = meta_kernel_ref.meta_baseline_bias_fwd_bwd(*[torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1, go)])
    print("bwd_data rel err", float((g[0].cpu() - want[1]).abs().max() / want[1].abs().max()))
