"""Times the Meta-Kernel kernels per implementation (CUDA events, inputs resident, B=4)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rangedet_b200 import _lib, ops, synth
B, C, H, W = 4, 64, 64, 2656
dev = "cuda"
data = torch.from_numpy(synth.feature_map(B, C, seed=1)).to(dev)
coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).to(dev)
w0, b0, w1, b1 = [torch.from_numpy(p).to(dev) for p in synth.meta_mlp_params(seed=2)]
go = torch.randn(B, 9 * C, H, W, device=dev)
out = torch.empty(B, 9 * C, H, W, device=dev)
gd = torch.empty_like(data)
L = _lib.lib()
P, S = ops._p, ops._stream
def timeit(fn, reps=10):
    """Device time per call: `reps` launches captured in ONE CUDA graph (no Python/ctypes launch gaps),
    replayed 3x, timed with CUDA events on the replay stream."""
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
px = B * H * W
for impl in (1, 3):
    os.environ.pop("RD_MK_TC_DEBUG", None)
    ms = timeit(lambda: _lib.check(L.rd_meta_kernel_fwd(P(data), P(coord), P(w0), P(b0), P(w1), P(b1), P(out), B, C, H, W, impl, S()), "fwd"))
    res["fwd_impl%d" % impl] = {"ms": ms, "GBps": px * 2572 / ms / 1e6}
for impl in (1, 3):
    ms = timeit(lambda: _lib.check(L.rd_meta_kernel_bwd_data(P(go), P(coord), P(w0), P(b0), P(w1), P(b1), P(gd), B, C, H, W, impl, S()), "bwd_data"))
    res["bwd_data_impl%d" % impl] = {"ms": ms, "GBps": px * 2572 / ms / 1e6}
gws = [torch.empty(96, device=dev), torch.empty(32, device=dev), torch.empty(C * 32, device=dev), torch.empty(C, device=dev)]
ws = torch.empty(int(L.rd_meta_kernel_bwd_workspace_bytes(B, C, H, W)) // 4 + 1, device=dev)
for impl in (1, 3):
    ms = timeit(lambda: _lib.check(L.rd_meta_kernel_bwd_params(P(go), P(data), P(coord), P(w0), P(b0), P(w1), P(b1), P(gws[0]), P(gws[1]), P(gws[2]), P(gws[3]), P(ws), ctypes.c_size_t(ws.numel() * 4), B, C, H, W, impl, S()), "bwd_params"))
    res["bwd_params_impl%d" % impl] = {"ms": ms, "GBps": px * 2572 / ms / 1e6}
for dbg in sys.argv[1:]:
    os.environ["RD_MK_TC_DEBUG"] = dbg
    res["fwd_impl2_dbg" + dbg] = {"ms": timeit(lambda: _lib.check(L.rd_meta_kernel_fwd(P(data), P(coord), P(w0), P(b0), P(w1), P(b1), P(out), B, C, H, W, 2, S()), "fwd"))}
# convolutions: the two dominant shapes of the DLA backbone / head
for (n, cin, cout, h, w) in [(4, 64, 64, 64, 2656), (4, 128, 128, 64, 664), (4, 128, 128, 64, 2656)]:
    x = ops.to_nhwc_padded(torch.randn(n, cin, h, w, device=dev))
    wt = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    y = torch.zeros((n, h + 2, w + 2, cout), device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.conv2d_nhwc(x, wt, None, None, relu=True, out=y))
    fl = 2.0 * n * h * w * cin * cout * 9
    res["conv3x3_%d_%d_w%d" % (cin, cout, w)] = {"ms": ms, "TFLOPs": fl / ms / 1e9}
print(json.dumps(res, indent=1))
