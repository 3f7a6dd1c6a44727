"""GPU parity tests (run on the B200 with `-m gpu`): every CUDA path, called through the C-ABI
(rangedet_b200.ops -> ctypes -> librangedet_b200.so), against the CPU oracle and the committed golden
vectors.  Tolerances (BASELINE.json north_star): fp32 results within 1e-3 relative (normwise, see
conftest.rel_err; tighter bounds asserted where the implementation allows), NMS keep indices
bit-exact.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, rel_err
from rangedet_b200 import synth

pytestmark = pytest.mark.gpu

REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(**kw):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(kw) + "\n")


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rangedet_b200 import ops as o
    return o


# ---------------------------------------------------------------------------------------------
# tcgen05 descriptor self-test
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mn_major", [False, True])
@pytest.mark.parametrize("n,k", [(64, 16), (64, 32), (64, 96), (16, 16), (128, 64), (256, 128), (80, 128)])
def test_tc_probe_gemm(ops, n, k, mn_major):
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g)
    b = torch.randn(n, k, generator=g)
    want = a.bfloat16().float() @ b.bfloat16().float().t()
    got = ops.tc_probe_gemm(a.cuda(), b.cuda(), mn_major=mn_major).cpu()
    err = rel_err(got.numpy(), want.numpy())
    report(test="tc_probe", n=n, k=k, mn_major=mn_major, rel_err=err)
    assert err < 1e-5, err


@pytest.mark.parametrize("box_w,c0,c1,c2", [(128, 0, 0, 0), (128, 4, 2, 0), (128, 200, 2, 0), (136, 132, 2, 0),
                                            (128, 0, 5, 0), (128, 0, 2, 64), (64, 60, 4, 64)])
def test_tma_probe(ops, box_w, c0, c1, c2):
    """TMA tile-mode rules this library relies on (probed on B200, scripts/tma_case.py): boxes may hang
    over the HIGH side of any dimension (zero fill on load, dropped on store); coordinates must be
    non-negative and the inner one 16-byte aligned (c0 = 1, -4, -128 and c1 = -1 all fault with
    'illegal instruction', which is why the Meta-Kernel loads start at max(w0 - 4, 0))."""
    C, H, W = 128, 5, 264
    src = torch.randn(C, H, W, device="cuda")
    tile, back = ops.tma_probe(src, box_w, c0, c1, c2)
    want = torch.zeros(64, box_w, device="cuda")
    want_back = torch.zeros_like(src)
    if 0 <= c1 < H:
        lo, hi = max(c0, 0), min(c0 + box_w, W)
        want[:, lo - c0:hi - c0] = src[c2:c2 + 64, c1, lo:hi]
        want_back[c2:c2 + 64, c1, lo:hi] = src[c2:c2 + 64, c1, lo:hi]
    assert torch.equal(tile, want)          # out-of-bound elements are zero-filled
    assert torch.equal(back, want_back)     # out-of-bound parts of a store are dropped


# ---------------------------------------------------------------------------------------------
# decode
# ---------------------------------------------------------------------------------------------
def test_decode_golden_and_oracle(ops, orc):
    g = golden("decode.npz")
    got = ops.decode_3d_bbox(cu(g["delta"]), cu(g["pc"])).cpu().numpy()
    np.testing.assert_allclose(got, g["out"], rtol=1e-4, atol=1e-4)
    got = ops.decode_3d_bbox(cu(g["delta_bin"]), cu(g["pc"]), is_bin=True).cpu().numpy()
    np.testing.assert_allclose(got, g["out_bin"], rtol=1e-4, atol=1e-4)
    # ragged size (not a multiple of the block), level-0 sized batch
    d, pc = synth.decode_inputs(2, 169984 // 8 + 13, seed=6)
    got = ops.decode_3d_bbox(cu(d), cu(pc)).cpu().numpy()
    want = orc.decode_3d_bbox(d, pc)
    err = rel_err(got, want)
    report(test="decode", rel_err=err)
    assert err < 1e-5
    # empty
    assert ops.decode_3d_bbox(torch.zeros(1, 0, 8).cuda(), torch.zeros(1, 0, 3).cuda()).shape == (1, 0, 10)
    with pytest.raises(ValueError):
        ops.decode_3d_bbox(torch.zeros(1, 4, 6).cuda(), torch.zeros(1, 4, 3).cuda())


# ---------------------------------------------------------------------------------------------
# rotated IoU
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t", [8, 5, 7])
def test_rotated_iou_golden(ops, t):
    g = golden("rotated_iou.npz")
    got = ops.rotated_iou(cu(g["a%d" % t]), cu(g["b%d" % t])).cpu().numpy()
    want = g["iou%d" % t]
    assert not np.isnan(got).any()
    diff = np.abs(got - want).max()
    report(test="rotated_iou_golden", box_type=t, max_abs_diff=float(diff), n_pos=int((want > 0).sum()))
    assert diff < 1e-3  # spec tolerance; typical 1e-6 (atan2f last-ulp sort swaps only)
    assert np.mean(np.abs(got - want) > 1e-5) < 1e-3


def test_rotated_iou_large_vs_oracle(ops, orc):
    c8 = synth.boxes7_to_corners10(synth.boxes7(6000, seed=21, clustered=True))[:, :8]
    gt = np.concatenate([c8[:100] + np.float32(0.11), synth.gt_boxes8(1, 0, 100)[0]], 0)  # incl. padded GT
    got = ops.rotated_iou(cu(c8), cu(gt)).cpu().numpy()
    want = orc.rotated_iou(c8, gt)
    assert np.abs(got - want).max() < 1e-3
    assert (want > 0.3).sum() > 100
    assert ops.rotated_iou(torch.zeros(0, 8).cuda(), cu(gt)).shape == (0, 200)


def test_batch_rotated_iou(ops, orc):
    B, N = 2, 20000
    gt = synth.gt_boxes8(B, n_real=50, n_total=200, seed=3)
    prop = np.zeros((B, N, 10), np.float32)
    for b in range(B):
        p = synth.boxes7_to_corners10(synth.boxes7(N, seed=40 + b, clustered=True))
        p[:50, :8] = gt[b, :50] + np.float32(0.1)
        prop[b] = p
    got = ops.batch_rotated_iou(cu(prop), cu(gt), "bev").cpu().numpy()
    want = orc.batch_rotated_iou_max(prop, gt, "bev")
    diff = np.abs(got - want).max()
    report(test="batch_rotated_iou_bev", max_abs_diff=float(diff), n_pos=int((want > 0).sum()))
    assert diff < 1e-3 and (want[:, :50] > 0.5).all()
    # '3d' mode
    gt7 = np.stack([synth.boxes7(200, seed=60 + b) for b in range(B)], 0)
    got = ops.batch_rotated_iou(cu(prop[:, :4000]), cu(gt7), "3d").cpu().numpy()
    want = orc.batch_rotated_iou_max(prop[:, :4000], gt7, "3d")
    assert np.abs(got - want).max() < 1e-3
    with pytest.raises(ValueError):
        ops.batch_rotated_iou(cu(prop), cu(gt), "xyz")


# ---------------------------------------------------------------------------------------------
# weighted NMS: keep indices bit-exact
# ---------------------------------------------------------------------------------------------
def _check_wnms(ops, dets, want_out, want_keep, is3d, tag):
    out, keep = ops.wnms_4c_device(cu(dets), 0.1, 0.5, is3d, 100)
    keep = keep.cpu().numpy()
    out = out.cpu().numpy()
    same = keep.shape == want_keep.shape and np.array_equal(keep, want_keep)
    report(test="wnms", tag=tag, is3d=bool(is3d), n=int(dets.shape[0]), K=int(len(want_keep)), K_gpu=int(len(keep)),
           keep_bit_exact=bool(same))
    assert same, "%s: keep indices differ (K %d vs %d)" % (tag, len(keep), len(want_keep))
    eq = (out == want_out) | (np.isnan(out) & np.isnan(want_out))
    if not eq.all():
        rows = np.where(~eq.all(1))[0]
        report(test="wnms_mismatch", tag=tag, n_rows=int(len(rows)), rows=rows[:5].tolist(),
               got=out[rows[:3]].tolist(), want=want_out[rows[:3]].tolist())
    assert eq.all(), "%s: merged boxes differ in %d rows" % (tag, int((~eq.all(1)).sum()))


@pytest.mark.parametrize("tag", ["clustered", "uniform"])
def test_wnms_golden(ops, tag):
    g = golden("wnms.npz")
    _check_wnms(ops, g[tag + "_dets"], g[tag + "_out"], g[tag + "_keep"], False, tag)
    _check_wnms(ops, g[tag + "_dets"], g[tag + "_out3d"], g[tag + "_keep3d"], True, tag + "_3d")


def test_wnms_vs_oracle_large(ops, orc):
    for n, cl, seed in [(20000, True, 1), (20000, False, 2), (100000, True, 0)]:
        dets = synth.wnms_dets(n, seed=seed, clustered=cl)
        wo, wk = orc.wnms_4c(dets, 0.1, 0.5, False, 100)
        _check_wnms(ops, dets, wo, wk, False, "n%d_%s" % (n, "clustered" if cl else "uniform"))


def test_wnms_parallel_path_3d_thresholds_and_overflow_fallback(ops, orc):
    """n >= 2048 takes the parallel path (adjacency of all boxes, then the light sequential walk); a cloud so dense that it
    overflows the candidate reservation (320 pairs per box on average) must fall back to the sequential scan; both bit-exact,
    also with the 3-D overlap, other thresholds and hash scales."""
    dets = synth.wnms_dets(6000, seed=21, clustered=True)
    for th, tv, is3d, hs in [(0.1, 0.5, True, 100), (0.3, 0.7, False, 100), (0.05, 0.2, False, 13)]:
        wo, wk = orc.wnms_4c(dets, th, tv, is3d, hs)
        go, gk = ops.wnms_4c_device(cu(dets), th, tv, is3d, hs)
        assert np.array_equal(gk.cpu().numpy(), wk) and np.array_equal(go.cpu().numpy(), wo, equal_nan=True), (th, tv, is3d, hs)
    # dense: 4000 boxes around 3 centres -> ~1300 mutually near boxes each, far more than 320 candidates per box
    rng = np.random.default_rng(5)
    b7 = synth.boxes7(4000, seed=6, clustered=False)
    centres = rng.uniform(-40, 40, (3, 2)).astype(np.float32)
    b7[:, :2] = centres[rng.integers(0, 3, 4000)] + rng.normal(0, 0.8, (4000, 2)).astype(np.float32)
    dense = synth.corners10_to_dets12(synth.boxes7_to_corners10(b7).astype(np.float32), synth.distinct_scores(4000, seed=7))
    wo, wk = orc.wnms_4c(dense, 0.1, 0.5, False, 100)
    go, gk = ops.wnms_4c_device(cu(dense), 0.1, 0.5, False, 100)
    assert np.array_equal(gk.cpu().numpy(), wk) and np.array_equal(go.cpu().numpy(), wo, equal_nan=True)


def test_wnms_edge_cases_and_plugin_surface(ops, orc):
    from rangedet_b200 import processing_cxx
    assert processing_cxx.wnms_4c(np.zeros((0, 12), np.float32), 0.1, 0.5, False, 100) == ([], [])
    one = synth.wnms_dets(1, seed=1)
    d, k = processing_cxx.wnms_4c(one, 0.1, 0.5, False, 100)
    assert isinstance(d, list) and isinstance(k, list) and k == [0] and len(d) == 12
    dets = synth.wnms_dets(3000, seed=9, clustered=True)
    d, k = processing_cxx.wnms_4c(dets, 0.1, 0.5, False, 100)
    wo, wk = orc.wnms_4c(dets, 0.1, 0.5, False, 100)
    assert k == wk.tolist()
    assert np.array_equal(np.array(d, np.float32).reshape(-1, 12), wo, equal_nan=True)
    # idempotence-style property: NMS of the kept ORIGINAL boxes keeps all of them
    kept = dets[np.array(k)]
    _, k2 = processing_cxx.wnms_4c(kept, 0.1, 0.5, False, 100)
    assert sorted(k2) == list(range(len(k)))
    # other thresholds / hash scale
    for th, tv, hs in [(0.3, 0.7, 100), (0.1, 0.5, 7)]:
        wo, wk = orc.wnms_4c(dets, th, tv, False, hs)
        go, gk = ops.wnms_4c_device(cu(dets), th, tv, False, hs)
        assert np.array_equal(gk.cpu().numpy(), wk) and np.array_equal(go.cpu().numpy(), wo, equal_nan=True)


# ---------------------------------------------------------------------------------------------
# NMS3D (hard NMS): keep indices bit-exact against the restated reference kernels
# ---------------------------------------------------------------------------------------------
def test_nms3d(ops, orc):
    for n, cl, thr, mk in [(3000, True, 0.1, 500), (20000, False, 0.25, 4096), (50000, True, 0.1, 1000)]:
        b = np.stack([synth.boxes7_to_corners10(synth.boxes7(n, seed=70 + i, clustered=cl)) for i in range(2)], 0)
        wk, wb = orc.nms3d(b, thr, mk, False)
        gk, gb = ops.nms3d(cu(b), thr, mk, False)
        same = np.array_equal(gk.cpu().numpy(), wk)
        report(test="nms3d", n=n, clustered=cl, kept=int((wk >= 0).sum()), keep_bit_exact=bool(same))
        assert same and np.array_equal(gb.cpu().numpy(), wb)
    wk, wb = orc.nms3d(b[:, :4000], 0.3, 300, True)
    gk, gb = ops.nms3d(cu(b[:, :4000]), 0.3, 300, True)
    assert np.array_equal(gk.cpu().numpy(), wk) and np.array_equal(gb.cpu().numpy(), wb)
    k0, b0 = ops.nms3d(torch.zeros(1, 0, 10).cuda(), 0.1, 8)
    assert (k0.cpu().numpy() == -1).all() and not b0.any()


# ---------------------------------------------------------------------------------------------
# Meta-Kernel
# ---------------------------------------------------------------------------------------------
def _mk_inputs(B, C, H, W, wpad, seed):
    coord = synth.range_image_coords(B, seed=seed, h=H, w=W, w_pad=wpad)
    data = synth.feature_map(B, C, seed=seed + 1, h=H, w=W, w_pad=wpad)
    params = synth.meta_mlp_params(seed=seed + 2, out_channels=C)
    return data, coord, params


@pytest.mark.parametrize("impl", [1, 3])
def test_meta_kernel_fwd_golden(ops, impl):
    g = golden("meta_kernel.npz")
    got = ops.meta_kernel_forward(cu(g["data"]), cu(g["coord"]), cu(g["w0"]), cu(g["b0"]), cu(g["w1"]), cu(g["b1"]),
                                  impl=impl).cpu().numpy()
    err = rel_err(got, g["out"])
    report(test="meta_fwd_golden", impl=impl, rel_err=err)
    assert err < (1e-5 if impl == 1 else 2e-4), err


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("shape", [(2, 64, 16, 300, 304), (1, 64, 3, 129, 132), (1, 64, 64, 2650, 2656)])
def test_meta_kernel_fwd_vs_oracle(ops, impl, shape):
    from oracle import meta_kernel_ref
    B, C, H, W, wpad = shape
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=10)
    want = meta_kernel_ref.meta_baseline_bias(*[torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1)]).numpy()
    got = ops.meta_kernel_forward(cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), impl=impl).cpu().numpy()
    err = rel_err(got, want)
    # element-wise: |diff| <= 1e-3 * (|want| + rms)   (the spec's 1e-3 rel with an absolute floor)
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2)))
    bad = np.abs(got - want) > 1e-3 * (np.abs(want) + rms)
    report(test="meta_fwd", impl=impl, shape=list(shape), rel_err=err, n_bad=int(bad.sum()))
    assert err < (1e-5 if impl == 1 else 2e-4), err
    assert not bad.any()


def test_meta_kernel_small_channel_counts(ops):
    from oracle import meta_kernel_ref
    for C in (8, 32):
        data, coord, (w0, b0, w1, b1) = _mk_inputs(1, C, 5, 70, 72, seed=20)
        want = meta_kernel_ref.meta_baseline_bias(*[torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1)]).numpy()
        got = ops.meta_kernel_forward(cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), impl=1).cpu().numpy()
        assert rel_err(got, want) < 1e-5
    with pytest.raises(RuntimeError):
        ops.meta_kernel_forward(torch.zeros(1, 12, 4, 8).cuda(), torch.zeros(1, 3, 4, 8).cuda(), torch.zeros(32, 3).cuda(),
                                torch.zeros(32).cuda(), torch.zeros(12, 32).cuda(), torch.zeros(12).cuda())


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("shape", [(1, 64, 5, 28, 28), (2, 64, 16, 300, 304), (1, 32, 7, 130, 130),
                                   (1, 64, 64, 2650, 2656)])
def test_meta_kernel_bwd_vs_oracle(ops, shape, impl):
    from oracle import meta_kernel_ref
    B, C, H, W, wpad = shape
    if impl == 3 and (C != 64 or wpad % 4):
        pytest.skip("impl 3 is specialised for C == 64, W % 4 == 0")
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=30)
    go = np.random.default_rng(5).standard_normal((B, 9 * C, H, wpad)).astype(np.float32)
    tt = [torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1, go)]
    want = meta_kernel_ref.meta_baseline_bias_fwd_bwd(*tt)[1:]
    got = ops.meta_kernel_backward(cu(go), cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), impl=impl)
    for name, g_, w_ in zip(["grad_data", "grad_w0", "grad_b0", "grad_w1", "grad_b1"], got, want):
        err = rel_err(g_.cpu().numpy().reshape(-1), w_.numpy().reshape(-1))
        report(test="meta_bwd", impl=impl, shape=list(shape), grad=name, rel_err=err)
        assert err < (1e-4 if impl == 1 else 2e-4), (name, err)


@pytest.mark.parametrize("B,N,K", [(1, 297472, 50000), (3, 5000, 700), (2, 17, 17), (1, 1, 1)])
def test_get_sorted_foreground_vs_oracle(ops, B, N, K):
    """Bit-exact (index work): distinct scores, a 60 % zero mask (massive ties at 0 incl. -0.0 from negative
    logits), duplicated scores."""
    from oracle import sorted_fg_ref
    rng = np.random.default_rng(B * 1000 + N)
    score = rng.standard_normal((B, N)).astype(np.float32)
    score[:, ::7] = np.round(score[:, ::7], 1)       # many exact duplicates
    mask = (rng.random((B, N)) > 0.6).astype(np.float32)
    delta = rng.standard_normal((B, N, 8)).astype(np.float32)
    pc = rng.standard_normal((B, N, 3)).astype(np.float32)
    want = sorted_fg_ref.get_sorted_foreground(score, delta, pc, mask, K)
    got = ops.get_sorted_foreground(cu(score), cu(delta), cu(pc), cu(mask), str(K))   # kwargs arrive as strings
    for g_, w_, name in zip(got, want, ("score", "bbox_delta", "pc")):
        assert np.array_equal(g_.cpu().numpy(), w_), name
    s = got[0].cpu().numpy()
    assert (np.diff(s, axis=1) <= 0).all()           # descending
    report(test="get_sorted_foreground", B=B, N=N, K=K, exact=True)


def test_get_sorted_foreground_errors(ops):
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError):
        ops.get_sorted_foreground(z(1, 10), z(1, 10, 8), z(1, 10, 3), z(1, 10), 11)   # num_fgs > N (:66)
    with pytest.raises(ValueError):
        ops.get_sorted_foreground(z(1, 10), z(1, 9, 8), z(1, 10, 3), z(1, 10), 5)


def test_meta_kernel_host_pipeline(ops):
    """Host-buffer call (pinned tensors, copy/compute pipeline) == device-resident call; twice, to cover
    slot reuse across calls."""
    B, C, H, W, wpad = 3, 64, 16, 300, 304
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=40)
    go = np.random.default_rng(6).standard_normal((B, 9 * C, H, wpad)).astype(np.float32)
    dw = [cu(x) for x in (w0, b0, w1, b1)]
    want_out = ops.meta_kernel_forward(cu(data), cu(coord), *dw)
    want = ops.meta_kernel_backward(cu(go), cu(data), cu(coord), *dw)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_data, h_coord, h_go = pin(data), pin(coord), pin(go)
    h_out, h_gd = torch.empty(B, 9 * C, H, wpad).pin_memory(), torch.empty(B, C, H, wpad).pin_memory()
    h_gp = torch.empty(96 + 32 + C * 32 + C).pin_memory()
    pipe = ops.MetaKernelHostPipeline(C, H, wpad, "cuda")
    for _ in range(2):
        h_out.zero_(); h_gd.zero_(); h_gp.zero_()
        pipe(h_data, h_coord, h_go, *dw, h_out, h_gd, h_gp)
        pipe.wait()
        torch.cuda.synchronize()
        assert torch.equal(h_out, want_out.cpu())          # same kernels, per frame: bit-identical
        assert torch.equal(h_gd, want[0].cpu())
        flat = torch.cat([g.reshape(-1) for g in want[1:]]).cpu()
        assert rel_err(h_gp.numpy(), flat.numpy()) < 1e-5   # summed per frame instead of per CTA row
    with pytest.raises(ValueError):
        pipe(torch.from_numpy(data), h_coord, h_go, *dw, h_out, h_gd, h_gp)  # not pinned


def test_meta_kernel_autograd_and_properties(ops):
    """Size-independent properties at the full BASELINE size: exact linearity in data (scaling by a
    power of two is exact in fp32), zero data -> zero output, determinism, autograd wiring."""
    B, C, H, W, wpad = 1, 64, 64, 2650, 2656
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=40)
    args = [cu(x) for x in (coord, w0, b0, w1, b1)]
    d = cu(data)
    for impl in (1, 3):
        o1 = ops.meta_kernel_forward(d, *args, impl=impl)
        o2 = ops.meta_kernel_forward(d * 2, *args, impl=impl)
        assert torch.equal(o2, o1 * 2)
        assert torch.equal(ops.meta_kernel_forward(d, *args, impl=impl), o1)  # deterministic
        assert not ops.meta_kernel_forward(torch.zeros_like(d), *args, impl=impl).any()
        assert not o1[..., W + 1:].any()  # columns whose whole 3x3 window has zero features stay zero
        del o1, o2
    dd = d[:, :, :8, :256].clone().requires_grad_(True)
    ps = [p.clone().requires_grad_(True) for p in args[1:]]
    out = ops.meta_kernel(dd, args[0][:, :, :8, :256].contiguous(), *ps)
    out.square().sum().backward()
    assert dd.grad is not None and all(p.grad is not None and torch.isfinite(p.grad).all() for p in ps)
    g1 = [p.grad.clone() for p in ps]
    dd.grad = None
    for p in ps:
        p.grad = None
    ops.meta_kernel(dd, args[0][:, :, :8, :256].contiguous(), *ps).square().sum().backward()
    assert all(torch.equal(a, p.grad) for a, p in zip(g1, ps))  # deterministic reduction


def test_meta_kernel_fused_nhwc_output(ops):
    """MODE 2: Meta-Kernel + per-channel scale/shift + ReLU -> haloed NHWC bf16, tap-major channels."""
    from oracle import meta_kernel_ref
    B, C, H, W, wpad = 2, 64, 6, 300, 304
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=50)
    g = torch.Generator().manual_seed(3)
    scale, shift = torch.rand(9 * C, generator=g) + 0.5, torch.randn(9 * C, generator=g) * 0.2
    ref = meta_kernel_ref.meta_baseline_bias(*[torch.from_numpy(x) for x in (data, coord, w0, b0, w1, b1)])
    want = (ref * scale[None, :, None, None] + shift[None, :, None, None]).relu()          # (B, c*9+k, H, W)
    yp = ops.meta_kernel_forward_nhwc(cu(data), cu(coord), cu(w0), cu(b0), cu(w1), cu(b1), scale.cuda(), shift.cuda())
    assert not yp[:, 0].any() and not yp[:, -1].any() and not yp[:, :, 0].any() and not yp[:, :, -1].any()
    got = yp[:, 1:-1, 1:-1, :].float().cpu().reshape(B, H, wpad, 9, C).permute(0, 4, 3, 1, 2).reshape(B, 9 * C, H, wpad)
    err = (got - want).abs()
    assert bool((err <= 2.0 ** -7 * want.abs() + 1e-2).all()), float(err.max())
    report(test="meta_fwd_nhwc_fused", rel_err=rel_err(got.numpy(), want.numpy()))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 6, 300, 304), (1, 64, 2650, 2656)])
def test_meta_kernel_backward_from_nhwc_gradient(ops, dtype, shape):
    """The training graph's Meta-Kernel backward reads the haloed NHWC tap-major gradient (2-byte storage) directly:
    same arithmetic on the widened values as the (B,9C,H,W) fp32 op-boundary path, so every output is bit-identical to
    meta_kernel_backward fed the converted tensor (which the oracle tests above pin)."""
    B, H, W, wpad = shape
    C = 64
    data, coord, (w0, b0, w1, b1) = _mk_inputs(B, C, H, W, wpad, seed=77)
    g = torch.Generator(device="cuda").manual_seed(5)
    go_pad = torch.zeros((B, H + 2, wpad + 2, 9 * C), device="cuda", dtype=dtype)
    go_pad[:, 1:-1, 1:-1] = torch.randn((B, H, wpad, 9 * C), device="cuda", generator=g).to(dtype)
    args = [cu(x) for x in (data, coord, w0, b0, w1, b1)]
    want = ops.meta_kernel_backward(ops.nhwc_to_nchw(go_pad, tap_major=True), *args, impl=3)
    got = ops.meta_kernel_backward_nhwc(go_pad, *args)
    for name, a, b in zip(("grad_data", "grad_w0", "grad_b0", "grad_w1", "grad_b1"), got, want):
        assert torch.equal(a, b), (name, float((a - b).abs().max()))
    only_p = ops.meta_kernel_backward_nhwc(go_pad, *args, need_data_grad=False)
    assert only_p[0] is None and all(torch.equal(a, b) for a, b in zip(only_p[1:], want[1:]))
    report(test="meta_bwd_nhwc", dtype=str(dtype), shape=list(shape), bit_identical=True)


def test_meta_kernel_class_surface(ops):
    from rangedet_b200.meta_kernel import MetaKernel
    mk = MetaKernel(num_batch=1, feat_height=8, feat_width=64, fp16=False)
    data = torch.randn(1, 64, 8, 64, device="cuda")
    coord = torch.randn(1, 3, 8, 64, device="cuda")
    out = mk.meta_baseline_bias(name="res1_unit2", data=data, coord_data=coord, data_channels=64, coord_channels=3,
                                channel_list=[32, 64], norm=None, conv1_filter=64, kernel_size=3)
    assert out.shape == (1, 576, 8, 64)
    assert sorted(mk.params) == ["res1_unit2_64_mlp0_bias", "res1_unit2_64_mlp0_weight",
                                 "res1_unit2_64_mlp1_bias", "res1_unit2_64_mlp1_weight"]


# ---------------------------------------------------------------------------------------------
# tcgen05 implicit-GEMM convolution (torch fp32 conv of the same bf16-rounded operands is the
# reference for this floating-point kernel; outputs are bf16: tolerance 2^-7 relative + small abs)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 64, 64, 5, 300, 3, 1), (2, 128, 128, 4, 200, 3, 1), (1, 64, 128, 3, 130, 3, 1),
                                   (1, 128, 64, 3, 257, 3, 1), (2, 64, 64, 3, 128, 1, 1), (1, 576, 64, 2, 140, 1, 1),
                                   (1, 8, 64, 4, 96, 3, 1), (1, 64, 64, 4, 300, 3, 2), (2, 64, 128, 3, 132, 1, 2),
                                   (1, 128, 128, 3, 258, 3, 2), (1, 72, 128, 3, 100, 3, 1), (1, 64, 64, 64, 2656, 3, 1)])
def test_conv2d_nhwc_vs_torch(ops, shape):
    import torch.nn.functional as F
    N, Cin, Cout, H, W, ks, sw = shape
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / (Cin * ks * ks) ** 0.5).bfloat16().float()
    scale = torch.rand(Cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(Cout, device="cuda", generator=g) * 0.1
    res = torch.randn(N, Cout, H, W // sw, device="cuda", generator=g).bfloat16().float()
    cin_p = ((Cin + 63) // 64) * 64
    xp = ops.to_nhwc_padded(x, cin_p)
    wp = ops.pack_conv_weight(w, cin_p, Cout)
    for relu, use_res in [(False, False), (True, True)]:
        want = F.conv2d(x, w, padding=ks // 2, stride=(1, sw)) * scale[None, :, None, None] + shift[None, :, None, None]
        if use_res:
            want = want + res
        if relu:
            want = want.relu()
        yp = ops.conv2d_nhwc(xp, wp, scale, shift, relu=relu, residual_pad=ops.to_nhwc_padded(res) if use_res else None,
                             stride_w=sw)
        got = ops.from_nhwc_padded(yp)
        # halo must stay zero (the next layer's padding)
        assert not yp[:, 0].any() and not yp[:, -1].any() and not yp[:, :, 0].any() and not yp[:, :, -1].any()
        err = (got - want).abs()
        tol = 2.0 ** -7 * want.abs() + 2e-2
        report(test="conv2d", shape=list(shape), relu=relu, residual=use_res, max_abs_err=float(err.max()),
               rel_err=rel_err(got.cpu().numpy(), want.cpu().numpy()))
        assert bool((err <= tol).all()), float((err - tol).max())


@pytest.mark.parametrize("shape", [(1, 128, 128, 4, 166, 8), (2, 128, 64, 3, 140, 8), (1, 128, 64, 3, 200, 4),
                                   (2, 64, 64, 5, 130, 4), (1, 64, 64, 3, 40, 8)])
def test_deconv2d_nhwc_vs_torch(ops, shape):
    """agg_stage transposed convolutions (dla_backbone.py:116-127): (3,8)/(1,4)/(1,2) and (3,4)/(1,2)/(1,1),
    followed by BN -> ReLU -> + skip."""
    import torch.nn.functional as F
    N, Cin, Cout, H, W, kw = shape
    S, pad = kw // 2, kw // 4
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16().float()
    w = (torch.randn(Cin, Cout, 3, kw, device="cuda", generator=g) / (Cin * 3 * 2) ** 0.5).bfloat16().float()
    scale = torch.rand(Cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(Cout, device="cuda", generator=g) * 0.1
    skip = torch.randn(N, Cout, H, W * S, device="cuda", generator=g).bfloat16().float()
    want = F.conv_transpose2d(x, w, stride=(1, S), padding=(1, pad)) * scale[None, :, None, None] + shift[None, :, None, None]
    want = want.relu() + skip
    yp = ops.deconv2d_nhwc(ops.to_nhwc_padded(x), ops.pack_deconv_weight(w), scale, shift, relu=True,
                           residual_pad=ops.to_nhwc_padded(skip))
    got = ops.from_nhwc_padded(yp)
    assert got.shape == want.shape
    assert not yp[:, 0].any() and not yp[:, -1].any() and not yp[:, :, 0].any() and not yp[:, :, -1].any()
    err = (got - want).abs()
    tol = 2.0 ** -7 * want.abs() + 2e-2
    report(test="deconv2d", shape=list(shape), max_abs_err=float(err.max()), rel_err=rel_err(got.cpu().numpy(), want.cpu().numpy()))
    assert bool((err <= tol).all()), float((err - tol).max())


# ---------------------------------------------------------------------------------------------
# full DLA backbone + Meta-Kernel + RPN head forward (inference form) vs the torch restatement
# ---------------------------------------------------------------------------------------------
def test_dla_backbone_and_head_forward(ops):
    from oracle import dla_ref
    from rangedet_b200 import dla
    B, H, W = 1, 8, 512
    P = dla_ref.make_params(seed=0, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    data = torch.randn(B, 8, H, W, device="cuda", generator=g)
    coord = cu(synth.range_image_coords(B, seed=3, h=H, w=W - 6, w_pad=W))
    ref = dla_ref.Ref(P, bf16=True)
    want_feats = ref.backbone(data, coord)
    want_cls, want_reg = ref.head(want_feats)
    bb = dla.DLABackbone(P)
    feats = bb.get_rpn_feature(data, coord)
    got_feats = [ops.from_nhwc_padded(feats[0], 72), ops.from_nhwc_padded(feats[1]), ops.from_nhwc_padded(feats[2])]
    for lvl, (gf, wf) in enumerate(zip(got_feats, want_feats)):
        assert gf.shape == wf.shape
        e = rel_err(gf.cpu().numpy(), wf.cpu().numpy())
        report(test="dla_backbone", level=lvl, shape=list(wf.shape), rel_err=e)
        assert e < 3e-2, (lvl, e)   # bf16 activations through ~45 layers
    cls, reg = dla.RangeRpnHead(P).get_fpn_output(feats)
    # the CUDA-graph replay of the same forward gives the same numbers
    gf = dla.GraphedForward(P, B, H, W)
    gcls, greg = gf(data, coord)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(gcls, cls)) and all(torch.equal(a, b) for a, b in zip(greg, reg))
    for lvl in range(3):
        for name, gq, wq in (("cls", cls[lvl], want_cls[lvl]), ("reg", reg[lvl], want_reg[lvl])):
            assert gq.shape == wq.shape
            e = rel_err(gq.cpu().numpy(), wq.cpu().numpy())
            report(test="rpn_head", level=lvl, out=name, shape=list(wq.shape), rel_err=e)
            assert e < 5e-2, (lvl, name, e)
