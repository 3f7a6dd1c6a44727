"""Why does meta_ws_params_kernel take 0.82 ms inside bench.py but 0.57 ms in mk_tc_diag.py?  Times the same
C-ABI call eagerly / graph-replayed, with the diag's and the bench's input recipes."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import _lib, ops, synth
B, C, H, W = 4, 64, 64, 2656
dev = "cuda"
L = _lib.lib()
P, S = ops._p, ops._stream
w0, b0, w1, b1 = [torch.from_numpy(p).to(dev) for p in synth.meta_mlp_params(seed=2)]
gws = [torch.empty(96, device=dev), torch.empty(32, device=dev), torch.empty(C * 32, device=dev), torch.empty(C, device=dev)]
ws = torch.empty(int(L.rd_meta_kernel_bwd_workspace_bytes(B, C, H, W)) // 4 + 1, device=dev)

def call(go, data, coord):
    _lib.check(L.rd_meta_kernel_bwd_params(P(go), P(data), P(coord), P(w0), P(b0), P(w1), P(b1), P(gws[0]), P(gws[1]), P(gws[2]),
                                           P(gws[3]), P(ws), ctypes.c_size_t(ws.numel() * 4), B, C, H, W, 3, S()), "bwd_params")

def eager(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def graphed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (3 * reps)

res = {}
for label, dseed, cseed in [("diag_inputs", 1, 0), ("bench_inputs", 100, 200)]:
    data = torch.from_numpy(synth.feature_map(B, C, seed=dseed)).to(dev)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=cseed)).to(dev)
    for glabel, go in [("randn", torch.randn(B, 9 * C, H, W, device=dev)), ("zeros", torch.zeros(B, 9 * C, H, W, device=dev))]:
        res["%s/%s/eager" % (label, glabel)] = eager(lambda: call(go, data, coord))
        res["%s/%s/graph" % (label, glabel)] = graphed(lambda: call(go, data, coord))
# bench-like process state: forward + backward-data run in between
data = torch.from_numpy(synth.feature_map(B, C, seed=100)).to(dev)
coord = torch.from_numpy(synth.range_image_coords(B, seed=200)).to(dev)
go = torch.randn(B, 9 * C, H, W, device=dev)
def step():
    out = ops.meta_kernel_forward(data, coord, w0, b0, w1, b1)
    ops.meta_kernel_backward(go, data, coord, w0, b0, w1, b1)
res["full_step/eager"] = eager(step)
print(json.dumps(res, indent=1))
