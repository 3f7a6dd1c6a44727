"""Per-shape timing of every tensor-core call of one eager training step (each call bracketed by CUDA events and a
synchronize, so calls run alone): which conv / wgrad shapes carry the step, and at what TFLOP/s.
    python scripts/conv_shapes_profile.py [batch]"""
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rangedet_b200 import ops, synth, train  # noqa: E402
from rangedet_b200.model_params import make_params  # noqa: E402

REC = defaultdict(lambda: [0, 0.0, 0.0])
ON = [False]


def wrap(name, flops_fn):
    orig = getattr(ops, name)

    def f(*a, **k):
        if not ON[0]:
            return orig(*a, **k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        torch.cuda.synchronize()
        key, fl = flops_fn(*a, **k)
        rec = REC[(name,) + key]
        rec[0] += 1
        rec[1] += e0.elapsed_time(e1)
        rec[2] += fl
        return r

    setattr(ops, name, f)


def conv_fl(x, w, *a, **k):
    N, Hp, Wp, Ci = x.shape
    taps, Co, _ = w.shape
    s = k.get("stride_w", 1)
    Wo = (Wp - 2) // s
    return (Ci, Co, taps, Wp - 2, s), 2.0 * N * (Hp - 2) * Wo * Ci * Co * taps


def slice_fl(x, w, out, c_off, **k):
    N, Hp, Wp, Ci = x.shape
    taps, Co, _ = w.shape
    return (Ci, Co, taps, Wp - 2, 1), 2.0 * N * (Hp - 2) * (Wp - 2) * Ci * Co * taps


def deconv_fl(x, w, *a, **k):
    N, Hp, Wp, Ci = x.shape
    taps, Co, _ = w.shape
    return (Ci, Co, taps, Wp - 2, 0), 2.0 * N * (Hp - 2) * (Wp - 2) * Ci * Co * taps


def wgrad_fl(a, b, ksize, stride_w=1, out=None):
    N, Hp, Wp, CA = a.shape
    CB = b.shape[3]
    return (CA, CB, ksize * ksize, Wp - 2, stride_w), 2.0 * N * (Hp - 2) * (Wp - 2) * CA * CB * ksize * ksize


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H, W = 64, 2656
    wrap("conv2d_nhwc", conv_fl)
    wrap("conv2d_nhwc_slice", slice_fl)
    wrap("deconv2d_nhwc", deconv_fl)
    wrap("conv2d_wgrad", wgrad_fl)
    P = make_params(seed=0, device="cuda")
    tg = train.TrainGraph(P)
    g = torch.Generator(device="cuda").manual_seed(1)
    data = torch.randn((B, 8, H, W), device="cuda", generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0)).cuda()
    d_cls = [torch.randn((B, 1, H, W >> l), device="cuda", generator=g) * 1e-3 for l in range(3)]
    d_reg = [torch.randn((B, 8, H, W >> l), device="cuda", generator=g) * 1e-3 for l in range(3)]
    for it in range(2):
        ON[0] = it == 1
        tg.forward(data, coord)
        tg.backward(d_cls, d_reg)
    tot = sum(v[1] for v in REC.values())
    print("batch %d: tensor-core calls of one step, timed alone: %.2f ms total" % (B, tot))
    print("%-20s %5s %5s %4s %6s %2s %4s %9s %8s %8s" % ("op", "Cin/A", "Cout/B", "taps", "W_in", "s", "n", "total us", "avg us", "TFLOP/s"))
    for k, v in sorted(REC.items(), key=lambda kv: -kv[1][1]):
        print("%-20s %5d %5d %4d %6d %2d %4d %9.1f %8.1f %8.1f" % (k[0], k[1], k[2], k[3], k[4], k[5], v[0], v[1] * 1e3, v[1] / v[0] * 1e3, v[2] / v[1] / 1e9))


if __name__ == "__main__":
    main()
