"""BatchNorm passes on the narrow layers (B=2, 128 ch): TMA-staged streaming kernels vs the register-staged ones
(RD_BN_STREAM=0), CUDA-graph replay of 20 back-to-back calls (launch overheads as inside the step)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rangedet_b200 import ops
dev, DT = "cuda", torch.float16
g = torch.Generator(device=dev).manual_seed(0)
B, H, C = 2, 64, 128
def graph_time(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); gr.replay(); b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / (2 * n) * 1e3, 2)
for w in (166, 332, 664, 1328):
    z = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
    dy = ops.to_nhwc_padded(torch.randn((B, C, H, w), device=dev, generator=g), dtype=DT)
    coef = ops.bn_train_stats(z, torch.ones(C, device=dev), torch.zeros(C, device=dev))
    y, dz = torch.zeros_like(z), torch.zeros_like(z)
    dgb = torch.empty((2, C), device=dev)
    r = {"W": w, "stream": os.environ.get("RD_BN_STREAM", "1")}
    r["fwd_apply_us"] = graph_time(lambda: ops.bn_act_fwd(z, coef, relu=True, out=y))
    r["bwd_us"] = graph_time(lambda: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz, dgb_out=dgb))
    print(json.dumps(r), flush=True)
