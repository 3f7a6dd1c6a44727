"""cfg-4: full DLA backbone + Meta-Kernel + RPN head forward (inference form), bf16, 64x2656.
Reports frames/s and achieved TFLOP/s (1.114 TFLOP / frame forward)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import dla, synth, _lib, model_params
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 64, 2656
P = model_params.make_params(seed=0, device="cuda")
data = torch.randn(B, 8, H, W, device="cuda")
coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).cuda()
bb, head = dla.DLABackbone(P), dla.RangeRpnHead(P)
def step():
    return head.get_fpn_output(bb.get_rpn_feature(data, coord))
for _ in range(2): step()
torch.cuda.synchronize()
l0 = _lib.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 5
for _ in range(reps): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
flops = 1.114e12 * B
eager_ms = ms
# the same forward captured once in a CUDA graph (static buffers: the activation pool) and replayed
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
g.replay(); torch.cuda.synchronize()
a.record()
for _ in range(reps): g.replay()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(json.dumps({"eager_ms": eager_ms, "timing": "CUDA graph replay of the whole forward", "workload": "DLA backbone + Meta-Kernel + RPN head forward, bf16, B=%d, 64x2656" % B, "ms": ms,
                  "frames_per_s": B / ms * 1e3, "TFLOPs_algorithmic": flops / ms / 1e9,
                  "frac_of_bf16_peak_1710": flops / ms / 1e9 / 1710.1, "rd_kernel_launches_per_step": (_lib.launch_count() - l0) / reps}))
