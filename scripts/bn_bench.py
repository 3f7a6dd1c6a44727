"""Stand-alone timing of the BatchNorm training passes (CUDA events, inputs larger than L2):
    python scripts/bn_bench.py [C ...]      env: RD_BN_STREAM=0|1, RD_BN_UB=<bytes>
Prints achieved GB/s (algorithmic bytes: every tensor read / written once) per kernel group."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangedet_b200 import ops  # noqa: E402


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    N, H, W = 4, 64, 2656
    out = {"env": {k: os.environ.get(k) for k in ("RD_BN_STREAM", "RD_BN_UB")}, "shape": [N, H, W], "results": {}}
    for C in [int(c) for c in sys.argv[1:]] or [64, 128]:
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda: (torch.randn((N, H + 2, W + 2, C), device="cuda", generator=g)).to(torch.bfloat16)
        z, dy, res = mk(), mk(), mk()
        y, dz = torch.zeros_like(z), torch.zeros_like(z)
        gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        nbytes = N * H * W * C * 2
        coef = ops.bn_train_stats(z, gamma, beta)
        r = {}
        t = timeit(lambda: ops.bn_train_stats(z, gamma, beta))
        r["stats+finalize"] = {"ms": t, "GBps": nbytes / t / 1e6}
        t = timeit(lambda: ops.bn_act_fwd(z, coef, relu=True, out=y))
        r["fwd_apply"] = {"ms": t, "GBps": 2 * nbytes / t / 1e6}
        t = timeit(lambda: ops.bn_act_fwd(z, coef, relu=True, res_before=res, out=y))
        r["fwd_apply+res"] = {"ms": t, "GBps": 3 * nbytes / t / 1e6}
        t = timeit(lambda: ops.bn_act_bwd(dy, z, coef, 2, dz_out=dz))
        r["bwd reduce+finalize+apply (mask from z)"] = {"ms": t, "GBps": 5 * nbytes / t / 1e6}
        t = timeit(lambda: ops.bn_act_bwd(dy, z, coef, 1, y_mask=y, dz_out=dz))
        r["bwd reduce+finalize+apply (mask from y)"] = {"ms": t, "GBps": 7 * nbytes / t / 1e6}
        out["results"][str(C)] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
