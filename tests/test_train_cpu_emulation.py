"""Host logic of the training graph on a box WITHOUT a GPU: rangedet_b200.train.TrainGraph runs with `train.ops`
replaced by tests/fake_ops.py (a plain-torch emulation of the kernel API with the same layouts), and is compared with
the torch-autograd restatement oracle/dla_train_ref.py.  This checks the tape -- data-gradient weight layouts, W-stride-2
layers as transposed convolutions, phase-grouped deconvolution gradients, the tap-major Meta-Kernel unit, residual
gradient accumulation -- and the flat-mode gather maps; the kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest
import torch

import fake_ops
from rangedet_b200 import synth, train


@pytest.fixture()
def cpu_train(monkeypatch):
    monkeypatch.setattr(train, "ops", fake_ops)
    return train


def _case(use_meta, B=1, H=4, W=32, seed=0):
    from oracle import dla_ref
    P = dla_ref.make_params(seed=seed, device="cpu")
    g = torch.Generator().manual_seed(9)
    P["res1_unit2_conv1_weight"] = torch.randn((64, 64, 3, 3), generator=g) * 0.06
    for k, v in (("gamma", 1.0), ("beta", 0.0), ("moving_mean", 0.0), ("moving_var", 1.0)):
        P["res1_unit2_bn1_" + k] = torch.full((64,), v)
    data = torch.randn(B, 8, H, W, generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0, h=H, w=W - 2, w_pad=W))
    d_cls = [torch.randn(B, 1, H, W // s, generator=g) * 1e-2 for s in (1, 2, 4)]
    d_reg = [torch.randn(B, 8, H, W // s, generator=g) * 1e-2 for s in (1, 2, 4)]
    return P, data, coord, d_cls, d_reg


def _rms(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float(((a - b) ** 2).mean().sqrt() / max(float((b ** 2).mean().sqrt()), 1e-30))


@pytest.fixture()
def exact_train(monkeypatch):
    """float64 activations and operands everywhere: no storage rounding, so the tape must agree with autograd to
    fp32-level precision (the parameter-gradient buffers stay float32)."""
    monkeypatch.setattr(train, "ops", fake_ops)
    fake_ops.set_exact(True)
    pack = train.pack_operand
    monkeypatch.setattr(train, "pack_operand", lambda w, kind, ci_p, co_p, S=1, dtype=None: pack(w, kind, ci_p, co_p, S, torch.float64))

    def get(self, key, shape):
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.bufs[key] = torch.zeros(tuple(shape), dtype=torch.float64)
        return t
    monkeypatch.setattr(train._Pool, "get", get)
    yield train
    fake_ops.set_exact(False)


@pytest.mark.parametrize("use_meta", [False, True], ids=["nometa", "meta"])
def test_train_graph_tape_equals_autograd_without_rounding(exact_train, use_meta):
    """Every activation, every data gradient and all 279 parameter gradients of the tape against torch autograd on the
    restatement, in float64: a wrong weight layout, tap order, phase grouping or residual routing gives O(1) errors;
    what remains is fp32 rounding of the gradient buffers."""
    from oracle import dla_train_ref
    P, data, coord, d_cls, d_reg = _case(use_meta, B=2, H=8, W=64)
    P = {k: v.double() for k, v in P.items()}
    data, coord = data.double(), coord.double()
    d_cls, d_reg = [d.double() for d in d_cls], [d.double() for d in d_reg]
    tg = exact_train.TrainGraph({k: v.clone() for k, v in P.items()}, device="cpu", use_meta=use_meta)
    cls, reg = tg.forward(data, coord)
    grads = tg.backward(d_cls, d_reg)
    rc, rr, rg = dla_train_ref.TrainRef(P, bf16=False, use_meta=use_meta).forward_backward(data, coord, d_cls, d_reg)
    for a, b in zip(cls + reg, rc + rr):
        assert _rms(a, b) < 1e-9
    used = [k for k in rg if float(rg[k].abs().max()) > 0]
    assert set(used) <= set(grads) and len(used) > 250
    errs = {k: _rms(grads[k].reshape(rg[k].shape), rg[k]) for k in used}
    assert max(errs.values()) < 1e-5, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    k = "res2_unit1_bn1_moving_var"          # moving statistics follow MXNet's update
    assert not torch.equal(tg.P[k], P[k])


@pytest.mark.parametrize("use_meta", [False, True], ids=["nometa", "meta"])
def test_train_graph_tape_with_bf16_storage_stays_within_the_noise_floor(cpu_train, use_meta):
    """Kernel-like storage (bf16 activations / gradients): at random initialisation the 40-layer BN + ReLU graph is
    chaotic on small images, so the bound is the restatement's own sensitivity to 2e-3 rounding-level jitter."""
    from oracle import dla_train_ref
    P, data, coord, d_cls, d_reg = _case(use_meta, B=2, H=8, W=64)
    tg = cpu_train.TrainGraph({k: v.clone() for k, v in P.items()}, device="cpu", use_meta=use_meta)
    cls, reg = tg.forward(data, coord)
    grads = tg.backward(d_cls, d_reg)
    rc, rr, rg = dla_train_ref.TrainRef(P, bf16=True, use_meta=use_meta).forward_backward(data, coord, d_cls, d_reg)
    jc, jr, jg = dla_train_ref.TrainRef(P, bf16=True, use_meta=use_meta, jitter=2e-3).forward_backward(data, coord, d_cls, d_reg)
    assert max(_rms(a, b) for a, b in zip(cls + reg, rc + rr)) < max(_rms(a, b) for a, b in zip(jc + jr, rc + rr))
    used = [k for k in rg if float(rg[k].abs().max()) > 0]
    e = np.median([_rms(grads[k].reshape(rg[k].shape), rg[k]) for k in used])
    floor = np.median([_rms(jg[k], rg[k]) for k in used])
    assert e < floor, (e, floor)


def test_tape_fuses_the_backward_sums_and_the_sliced_data_gradient_where_it_should(cpu_train, monkeypatch):
    """Which layers take the fused paths is host logic: conv1 -> bn1 -> relu -> conv2 chains of the 128-channel stride-1 blocks and
    the head towers hand their BatchNorm-backward sums to the data-gradient conv above (TrainGraph.conv_bn, sole_consumer); the
    level-0 head towers compute the data gradient of the 64 agg3 channels of concat(data, agg3) only; switching the fusion off
    changes no gradient."""
    P, data, coord, d_cls, d_reg = _case(True, B=2, H=8, W=64)
    calls = {"bwdstats": [], "sums": 0, "slice": []}
    real_bwdstats, real_bn_bwd, real_conv = fake_ops.conv2d_nhwc_bwdstats, fake_ops.bn_act_bwd, fake_ops.conv2d_nhwc

    def bwdstats(x_pad, w_packed, bn_z_pad, bn_coef, bn_mask_mode, out=None, ws=None):
        calls["bwdstats"].append((tuple(bn_z_pad.shape), int(bn_mask_mode), int(w_packed.shape[1])))
        return real_bwdstats(x_pad, w_packed, bn_z_pad, bn_coef, bn_mask_mode, out=out, ws=ws)

    def bn_bwd(*a, **k):
        calls["sums"] += k.get("sums") is not None
        return real_bn_bwd(*a, **k)

    def conv(x_pad, w_packed, *a, **k):
        if w_packed.shape[0] == 9 and x_pad.shape[3] == 128 and w_packed.shape[1] == 64 and x_pad.shape[2] == 64 + 2:
            calls["slice"].append(tuple(w_packed.shape))
        return real_conv(x_pad, w_packed, *a, **k)

    monkeypatch.setattr(fake_ops, "conv2d_nhwc_bwdstats", bwdstats)
    monkeypatch.setattr(fake_ops, "bn_act_bwd", bn_bwd)
    monkeypatch.setattr(fake_ops, "conv2d_nhwc", conv)
    grads = {}
    for fuse in (True, False):
        tg = cpu_train.TrainGraph({k: v.clone() for k, v in P.items()}, device="cpu", use_meta=True)
        tg.fuse_bwd_sums = fuse
        tg.forward(data, coord)
        grads[fuse] = tg.backward(d_cls, d_reg)
        if fuse:
            n_fused = len(calls["bwdstats"])
            # 12 backbone blocks (the 128-channel stride-1 units of res2a / res2 / res3a / res3 and the aggregation stages) + 18 head layers
            assert n_fused == 30 and calls["sums"] == n_fused, (n_fused, calls["sums"])
            assert all(z[3] == 128 and mm == 2 and co == 128 for z, mm, co in calls["bwdstats"])
            assert len(calls["slice"]) == 2, calls["slice"]        # cls and reg tower of level 0: 128 -> 64 data gradient
        else:
            assert len(calls["bwdstats"]) == n_fused and calls["sums"] == n_fused      # nothing added with the fusion off
    for k in grads[True]:
        assert torch.equal(grads[True][k], grads[False][k]), k


def test_flat_mode_gathers_reproduce_the_per_tensor_path(cpu_train):
    P, data, coord, d_cls, d_reg = _case(True)
    names = sorted(k for k in P if not k.endswith(("_moving_mean", "_moving_var")))
    # eager reference
    tg0 = cpu_train.TrainGraph({k: v.clone() for k, v in P.items()}, device="cpu")
    tg0.forward(data, coord)
    g0 = {k: v.clone() for k, v in tg0.backward(d_cls, d_reg).items()}
    packed0 = {k: v.clone() for k, v in tg0.packed.items()}
    # flat parameters / gradients like GraphedTrainStep builds them
    Pf = {k: v.clone() for k, v in P.items()}
    sizes = [Pf[k].numel() for k in names]
    flatP, flat_g = torch.empty(sum(sizes)), torch.zeros(sum(sizes))
    offsets, o = {}, 0
    for k, n in zip(names, sizes):
        offsets[k] = o
        flatP[o:o + n] = Pf[k].reshape(-1)
        Pf[k] = flatP[o:o + n].view(Pf[k].shape)
        o += n
    tg = cpu_train.TrainGraph(Pf, device="cpu")

    def run_step():
        tg.refresh()
        tg.forward(data, coord)
        tg.backward(d_cls, d_reg)

    run_step()
    tg.enable_flat(flatP, offsets, flat_g, run_step)
    assert tg.no_grad_params == sorted(k for k in names if k.startswith("res1_unit2_conv1") or k.startswith("res1_unit2_bn1"))
    run_step()                                                 # flat mode: one gather in, one gather out
    # backward in buckets (head | aggregation stages + res3 / res3a | res2 / res2a | res1): the same flat gradient, bucket
    # by bucket -- what the overlapped all-reduces of GraphedTrainStep exchange
    whole = flat_g.clone()
    ranges = tg.bucket_ranges()
    covered = sorted(r for rs in ranges for r in rs)
    assert covered[0][0] == 0 and covered[-1][1] == flat_g.numel() and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    assert [len(r) for r in ranges] == [1, 2, 1, 1]      # rpn_* | agg*, res3* | res2* | res1*
    tg.refresh()
    tg.forward(data, coord)
    flat_g.zero_()
    done = torch.zeros(flat_g.numel(), dtype=torch.bool)
    for k in range(tg.N_BUCKETS):
        tg.backward_bucket(k, d_cls, d_reg)
        for lo, hi in ranges[k]:
            done[lo:hi] = True
        assert torch.equal(flat_g[done], whole[done]) and not flat_g[~done].any(), k
    assert torch.equal(flat_g, whole)
    for key, want in packed0.items():                          # every bf16 operand re-packed by the gather
        assert torch.equal(tg.packed[key], want), key
    for k in names:
        want = g0[k].reshape(-1) if k in g0 else torch.zeros(Pf[k].numel())
        assert torch.equal(flat_g[offsets[k]:offsets[k] + Pf[k].numel()], want), k
    # the optimiser restatement on the flat buffers == per-tensor MXNet SGD
    wd = torch.cat([torch.full((Pf[k].numel(),), 1e-5 * train.wd_mult(k)) for k in names])
    mom = torch.zeros_like(flatP)
    before = {k: Pf[k].clone() for k in names}
    fake_ops.sgd_mom_update(flatP, flat_g, mom, wd, torch.tensor([0.05, 0.9, 1 / 128, 35.0]))
    Ps, ms = {k: before[k].clone() for k in names}, {}
    train.sgd_momentum_step(Ps, {k: g0.get(k, torch.zeros_like(before[k])) for k in names}, ms, lr=0.05, wd=1e-5,
                            clip_gradient=35.0, rescale_grad=1 / 128)
    for k in names:
        assert torch.allclose(Pf[k], Ps[k], rtol=1e-6, atol=1e-7), k


def test_whole_training_iteration_on_cpu(cpu_train):
    """GraphedTrainStep(capture=False): same flat buffers, gather maps, loss hand-off and optimiser call as the captured
    step, launched eagerly -- forward, fused-loss restatement, backward, SGD on the emulated kernels."""
    from oracle import dla_ref
    B, H, W = 1, 8, 64
    P = dla_ref.make_params(seed=0, device="cpu")
    step = cpu_train.GraphedTrainStep(P, B, H, W, lr=0.05, device="cpu", capture=False, overlap_wgrad=False)
    T = synth.rpn_targets(B, seed=7, n_vehicles=6, h=H, w=W - 6, w_pad=W)
    step.set_targets(T)
    g = torch.Generator().manual_seed(4)
    data = torch.randn(B, 8, H, W, generator=g)
    xyz = torch.from_numpy(T["pc_vehicle_frame_s1"]).reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    coord = xyz / torch.tensor([25.0, 25.0, 2.0]).view(1, 3, 1, 1)
    before = step.flatP.clone()
    hist = []
    for it in range(5):
        out = step.train_step(data, coord)
        hist.append(sum(float(o["reg_loss"].sum()) for o in out))
        assert np.isfinite(hist[-1]) and bool(torch.isfinite(step.flatP).all())
        if it == 0:
            dead = [k for k in step.names if float(step.gviews[k].abs().max()) == 0.0 and k not in step.tg.no_grad_params]
            assert not dead, dead[:5]
    assert min(hist[2:]) < hist[0], hist                    # the regression loss goes down
    assert float((step.flatP - before).abs().max()) > 1e-4
    assert all(P[k].data_ptr() >= step.flatP.data_ptr() for k in step.names)   # masters live in the flat buffer
    step.set_lr(0.0)
    w = step.flatP.clone()
    step.train_step(data, coord)
    assert torch.equal(step.flatP, w + step.flat_m)          # lr = 0: pure momentum step


def test_inference_graph_host_logic_on_cpu(monkeypatch):
    """rangedet_b200.dla (folded moving statistics, fused Meta-Kernel unit with tap-major aggregation weights, deconv
    skip connections, head) over the emulated kernels vs the restatement oracle/dla_ref.py with the same bf16 storage."""
    from oracle import dla_ref
    from rangedet_b200 import dla
    monkeypatch.setattr(dla, "ops", fake_ops)
    P = dla_ref.make_params(seed=0, device="cpu")
    g = torch.Generator().manual_seed(2)
    for k in P:                                    # non-trivial statistics so that the folding matters
        if k.endswith("_moving_mean"):
            P[k] = 0.1 * torch.randn(P[k].shape, generator=g)
        elif k.endswith("_moving_var"):
            P[k] = 1 + 0.3 * torch.rand(P[k].shape, generator=g)
        elif k.endswith("_gamma"):
            P[k] = 1 + 0.2 * torch.randn(P[k].shape, generator=g)
    B, H, W = 1, 8, 64
    data = torch.randn(B, 8, H, W, generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0, h=H, w=W - 2, w_pad=W))
    cls, reg = dla.RangeRpnHead(P, "cpu").get_fpn_output(dla.DLABackbone(P, "cpu").get_rpn_feature(data, coord))
    ref = dla_ref.Ref(P, bf16=True)
    rc, rr = ref.head(ref.backbone(data, coord))
    for a, b in zip(cls + reg, rc + rr):
        assert a.shape == b.shape and _rms(a, b) < 3e-2, _rms(a, b)


def test_inference_executor_equals_reference_graph_without_rounding(monkeypatch):
    """Forward (rangedet_b200.dla, folded moving statistics) + symbol._TestExecutor.predict over the emulated kernels
    in float64, against the reference's OWN inference graph executed eagerly (DLABackbone + get_fpn_output +
    get_fpn_prediction through oracle/mx_eager.py).  Skipped where /root/reference is absent."""
    from oracle import dla_ref, mx_eager, ref_graph
    if not mx_eager.available():
        pytest.skip("/root/reference not present")
    from rangedet_b200 import dla, symbol
    from test_symbol import shipped_config
    fake_ops.set_exact(True)
    try:
        monkeypatch.setattr(dla, "ops", fake_ops)
        monkeypatch.setattr(symbol, "ops", fake_ops)
        pk_c, pk_d = fake_ops.pack_conv_weight, fake_ops.pack_deconv_weight
        monkeypatch.setattr(fake_ops, "pack_conv_weight", lambda w, cin=None, cout=None, dtype=None: pk_c(w, cin, cout, torch.float64))
        monkeypatch.setattr(fake_ops, "pack_deconv_weight", lambda w, cin=None, cout=None, dtype=None: pk_d(w, cin, cout, torch.float64))

        def get(self, key, shape, device):
            t = self.bufs.get(key)
            if t is None or tuple(t.shape) != tuple(shape):
                t = self.bufs[key] = torch.zeros(tuple(shape), dtype=torch.float64)
            return t
        monkeypatch.setattr(dla._BufferPool, "get", get)

        def to_pad(x, channels=None, dtype=None):
            N, C, H, W = x.shape
            out = torch.zeros((N, H + 2, W + 2, channels or C), dtype=torch.float64)
            out[:, 1:H + 1, 1:W + 1, :C] = x.permute(0, 2, 3, 1)
            return out
        monkeypatch.setattr(fake_ops, "to_nhwc_padded", to_pad)
        monkeypatch.setattr(fake_ops, "from_nhwc_padded",
                            lambda y, channels=None: (y[:, 1:-1, 1:-1, :channels] if channels else y[:, 1:-1, 1:-1]).permute(0, 3, 1, 2).contiguous())
        B, H, W = 1, 16, 64
        P = dla_ref.make_params(seed=0, device="cpu")
        g = torch.Generator().manual_seed(3)
        data = torch.randn(B, 8, H, W, generator=g)
        T = synth.rpn_targets(B, seed=3, n_vehicles=5, h=H, w=W - 2, w_pad=W)
        xyz = torch.from_numpy(T["pc_vehicle_frame_s1"]).reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
        coord = xyz / torch.tensor([25.0, 25.0, 2.0]).view(1, 3, 1, 1)
        P.update(ref_graph.backbone_head(P, data, coord, training=True, bn_momentum=0.0)["moving"])   # "trained" statistics
        P64 = {k: v.double() for k, v in P.items()}
        pB, pR, _ = shipped_config(False, (H, W))
        ex = symbol._TestExecutor.__new__(symbol._TestExecutor)
        ex.sym = symbol.RangeRCNN(pR).get_test_symbol(symbol.DLABackbone(pB), symbol.RangeRpnHead(pR))
        ex.pre_n, ex.post_n, ex.nms_thr, ex.wnms = 300, 200, 0.2, True
        cls, reg = dla.RangeRpnHead(P64, "cpu").get_fpn_output(dla.DLABackbone(P64, "cpu").get_rpn_feature(data.double(), coord.double()))
        record = {k: torch.from_numpy(v) for k, v in T.items()}
        score, boxes, _ = ex.predict(cls, reg, record)
        r = ref_graph.backbone_head(P64, data.double(), coord.double(), training=False)
        for a, b in zip(cls + reg, r["cls"] + r["reg"]):
            assert _rms(a, b) < 1e-5            # folded scale / shift are fp32 in dla._Layer
        sc_r, box_r = ref_graph.fpn_prediction([c.float() for c in r["cls"]], [d.float() for d in r["reg"]],
                                               [T["pc_vehicle_frame_s%d" % s] for s in (1, 2, 4)],
                                               [T["range_image_mask_s%d" % s].reshape(B, -1) for s in (1, 2, 4)], 300)
        assert float((score - sc_r).abs().max()) < 1e-5
        d = torch.cdist(boxes[0].double(), box_r[0].double(), p=float("inf")).min(1).values     # near-ties may swap ranks
        assert float(d.max()) < 1e-2 and int(((boxes - box_r).abs().amax(-1) > 1e-2).sum()) < 15
    finally:
        fake_ops.set_exact(False)


def test_whole_train_step_equals_reference_graph_without_rounding(exact_train):
    """The strongest single check of the training path's LOGIC: GraphedTrainStep (tape + fused-loss hand-off + flat
    gradient collection) over the emulated kernels in float64, against the reference's OWN train graph
    (get_train_symbol: backbone, head, get_fpn_loss with MakeLoss) executed eagerly -- head outputs, the six loss
    tensors, and EVERY parameter gradient.  Skipped where /root/reference is absent."""
    from oracle import dla_ref, mx_eager, ref_graph
    if not mx_eager.available():
        pytest.skip("/root/reference not present")
    B, H, W = 1, 64, 32                                   # get_vfl_loss hard-codes 64 rows (builder.py:367)
    P = dla_ref.make_params(seed=0, device="cpu")
    g = torch.Generator().manual_seed(1)
    for k in P:
        if k.endswith("_gamma"):
            P[k] = 1 + 0.2 * torch.randn(P[k].shape, generator=g)
        elif k.endswith(("_beta", "_bias")):
            P[k] = 0.1 * torch.randn(P[k].shape, generator=g)
    data = torch.randn(B, 8, H, W, generator=g)
    T = synth.rpn_targets(B, seed=3, n_vehicles=6, h=H, w=W - 2, w_pad=W)
    xyz = torch.from_numpy(T["pc_vehicle_frame_s1"]).reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    coord = xyz / torch.tensor([25.0, 25.0, 2.0]).view(1, 3, 1, 1)
    P64 = {k: v.double() for k, v in P.items()}
    r = ref_graph.backbone_head(P64, data.double(), coord.double(), training=True, targets=T)      # fp32 graph: scale_loss_shift 1
    hyper = dict(exact_train.LOSS_HYPER, scale_loss_shift=1.0)
    step = exact_train.GraphedTrainStep({k: v.clone() for k, v in P64.items()}, B, H, W, lr=0.0, rescale_grad=1.0, device="cpu",
                                        capture=False, overlap_wgrad=False, loss_hyper=hyper)
    step.set_targets(T)
    out = step.train_step(data, coord)
    for a, b in zip(step.out[0] + step.out[1], r["cls"] + r["reg"]):
        assert _rms(a, b) < 1e-5                  # the flat master parameters are float32
    for lvl in range(3):
        assert _rms(out[lvl]["cls_loss"], r["cls_loss"][lvl]) < 1e-5 and _rms(out[lvl]["reg_loss"], r["reg_loss"][lvl]) < 1e-5
    used = [k for k in step.names if k in r["grads"]]
    assert len(used) > 270
    errs = {k: _rms(step.gviews[k], r["grads"][k].reshape(step.gviews[k].shape)) for k in used}
    # the regression towers see identical loss gradients; the IoU target behind the classification gradient is an fp32,
    # piecewise function of the regression outputs (loss.py:25-27: target > 0 vs == 0), so a handful of pixels flip
    # branch under the 1e-6 perturbation of the fp32 master weights -> 1e-3 on the classification side
    assert max(v for k, v in errs.items() if k.startswith("rpn_reg")) < 1e-4, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert max(errs.values()) < 5e-3, sorted(errs.items(), key=lambda kv: -kv[1])[:5]
