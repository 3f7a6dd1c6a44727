"""Times the tcgen05 Meta-Kernel forward under the diagnostic switches (RD_MK_TC_DEBUG)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import ops, synth
B, C = 4, 64
dev = "cuda"
data = torch.from_numpy(synth.feature_map(B, C, seed=1)).to(dev)
coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).to(dev)
ps = [torch.from_numpy(p).to(dev) for p in synth.meta_mlp_params(seed=2)]
res = {}
for dbg in sys.argv[1:] or ["0", "1", "2", "3", "4", "8", "11", "15"]:
    os.environ["RD_MK_TC_DEBUG"] = dbg
    for _ in range(3):
        ops.meta_kernel_forward(data, coord, *ps, impl=2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.meta_kernel_forward(data, coord, *ps, impl=2)
    b.record()
    torch.cuda.synchronize()
    res[dbg] = a.elapsed_time(b) / 10
print(json.dumps(res))
