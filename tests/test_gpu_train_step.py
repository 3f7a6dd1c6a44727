"""GPU tests of the captured training iteration (rangedet_b200.train.GraphedTrainStep, run with `-m gpu`):
flat-mode plumbing (one gather packs all operands, one collects all gradients, one SGD launch) against the
eager per-tensor path of the SAME kernels, and the optimiser kernel against the MXNet SGD restatement
(tools/train.py:306-319, 359-361).  All through the C-ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 8, 256


def _params(seed=0):
    from rangedet_b200.model_params import make_params
    P = make_params(seed=seed, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(9)   # the plain-conv variant of the Meta-Kernel unit (use_meta=False)
    P["res1_unit2_conv1_weight"] = torch.randn((64, 64, 3, 3), device="cuda", generator=g) * 0.06
    for k, v in (("gamma", 1.0), ("beta", 0.0), ("moving_mean", 0.0), ("moving_var", 1.0)):
        P["res1_unit2_bn1_" + k] = torch.full((64,), v, device="cuda")
    return P


def test_gather_kernels():
    from rangedet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    src = torch.randn(10007, device="cuda", generator=g)
    for n in (1, 2, 4097, 300001):
        idx = torch.randint(-1, src.numel(), (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
        want = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), device="cuda"))
        o32 = ops.gather_f32(src, idx, torch.empty(n, device="cuda"))
        assert torch.equal(o32, want)
        o16 = ops.gather_to_bf16(src, idx, torch.empty(n, device="cuda", dtype=torch.bfloat16))
        assert torch.equal(o16, want.to(torch.bfloat16))


def test_sgd_mom_update_matches_mxnet_restatement():
    from rangedet_b200 import ops, train
    g = torch.Generator(device="cuda").manual_seed(1)
    names = ["a_weight", "a_bias", "bn_gamma", "bn_beta"]
    shapes = [(64, 8, 3, 3), (64,), (64,), (64,)]
    P = {n: torch.randn(s, device="cuda", generator=g) for n, s in zip(names, shapes)}
    G = {n: torch.randn(s, device="cuda", generator=g) * 6000.0 for n, s in zip(names, shapes)}   # some elements clip at 35*128
    flat = lambda d: torch.cat([d[n].reshape(-1) for n in names])
    w, gr, m = flat(P), flat(G), torch.zeros(sum(P[n].numel() for n in names), device="cuda")
    wd = torch.cat([torch.full((P[n].numel(),), 1e-5 * train.wd_mult(n), device="cuda") for n in names])
    hyper = torch.tensor([0.05, 0.9, 1.0 / 128, 35.0], device="cuda")
    mom = {}
    for it in range(3):
        ops.sgd_mom_update(w, gr, m, wd, hyper)
        train.sgd_momentum_step(P, G, mom, lr=0.05, momentum=0.9, wd=1e-5, clip_gradient=35.0, rescale_grad=1.0 / 128)
        assert torch.allclose(w, flat(P), rtol=1e-6, atol=1e-6), it
        assert torch.allclose(m, flat(mom), rtol=1e-6, atol=1e-6), it
    assert float((gr.abs() / 128 > 35).float().mean()) > 0.01   # clipping was exercised
    assert train.wd_mult("x_bias") == 0.0 and train.wd_mult("x_beta") == 0.0 and train.wd_mult("x_gamma") == 1.0


@pytest.mark.parametrize("use_meta", [False, True], ids=["nometa", "meta"])
def test_graphed_step_matches_eager_path(use_meta):
    """Flat mode (CUDA-graph replay) must reproduce the eager TrainGraph: same kernels, only the plumbing differs
    -> operand copies bit-identical, gradients bit-identical, and one SGD step lands on the restated update."""
    from rangedet_b200 import synth, train
    B = 2
    g = torch.Generator(device="cuda").manual_seed(3)
    data = torch.randn((B, 8, H, W), device="cuda", generator=g)
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0, h=H, w=W - 6, w_pad=W)).cuda()
    d_cls = [torch.randn((B, 1, H, W // s), device="cuda", generator=g) * 1e-2 for s in (1, 2, 4)]
    d_reg = [torch.randn((B, 8, H, W // s), device="cuda", generator=g) * 1e-2 for s in (1, 2, 4)]
    # eager reference
    Pe = _params()
    tg = train.TrainGraph(Pe, use_meta=use_meta)
    cls_e, reg_e = tg.forward(data, coord)
    cls_e, reg_e = [t.clone() for t in cls_e], [t.clone() for t in reg_e]
    grads_e = {k: v.clone() for k, v in tg.backward(d_cls, d_reg).items()}
    # graphed, external gradients
    Pg = _params()
    P0 = {k: v.clone() for k, v in Pg.items()}
    step = train.GraphedTrainStep(Pg, B, H, W, lr=0.02, use_meta=use_meta, with_loss=False)
    for k in P0:   # construction (lr = 0 warm-up on zero inputs) must move neither the trainable parameters nor the
        assert torch.equal(Pg[k], P0[k]), k   # BatchNorm moving statistics (a loaded checkpoint survives bind())
    cls_g, reg_g = step.forward(data, coord)
    for a, b in zip(cls_e + reg_e, list(cls_g) + list(reg_g)):
        assert torch.equal(a, b)
    step.backward_update(d_cls, d_reg)
    torch.cuda.synchronize()
    unused = [k for k in step.names if k not in grads_e]   # the unit variant this graph does not contain
    assert unused == step.tg.no_grad_params and all(k.startswith("res1_unit2") for k in unused), unused[:5]
    for k in step.names:
        want = grads_e[k].reshape(step.gviews[k].shape) if k in grads_e else torch.zeros_like(step.gviews[k])
        assert torch.equal(step.gviews[k], want), k
    # the update: w1 = w0 + m1, m1 = -lr * (clip(g/128) + wd*wd_mult*w0)
    for k in step.names:
        ge = grads_e[k].reshape(P0[k].shape) if k in grads_e else torch.zeros_like(P0[k])
        gk = (ge / 128.0).clamp(-35, 35) + 1e-5 * train.wd_mult(k) * P0[k]
        want = P0[k] - 0.02 * gk
        assert torch.allclose(Pg[k], want, rtol=1e-5, atol=1e-7), k
    # second replay works on the updated weights (operands re-packed from the flat masters)
    cls2, _ = step.forward(data, coord)
    assert not torch.equal(cls2[0], cls_e[0])
    step.set_lr(0.0)
    before = step.flatP.clone()
    step.backward_update(d_cls, d_reg)
    torch.cuda.synchronize()
    assert torch.equal(step.flatP, before + step.flat_m)   # lr = 0: pure momentum step (set_lr reaches the captured graph)
    assert bool(torch.isfinite(step.flatP).all())


def test_graphed_step_with_rpn_loss_decreases_loss():
    """train_step() with the fused RPN loss on a synthetic roidb record: losses finite, gradients flow to every
    parameter, and a few SGD steps reduce the regression loss."""
    from rangedet_b200 import synth, train
    B = 2
    P = _params()
    step = train.GraphedTrainStep(P, B, H, W, lr=0.05, use_meta=True)
    T = synth.rpn_targets(B, seed=7, n_vehicles=6, h=H, w=W - 6, w_pad=W)
    step.set_targets(T)
    g = torch.Generator(device="cuda").manual_seed(4)
    data = torch.randn((B, 8, H, W), device="cuda", generator=g)
    xyz = torch.from_numpy(T["pc_vehicle_frame_s1"]).cuda().reshape(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    coord = xyz / torch.tensor([25.0, 25.0, 2.0], device="cuda").view(1, 3, 1, 1)
    hist = []
    for it in range(8):
        out = step.train_step(data, coord)
        torch.cuda.synchronize()
        tot = sum(float(o["reg_loss"].sum()) for o in out)
        cls = sum(float(o["cls_loss"].sum()) for o in out)
        assert np.isfinite(tot) and np.isfinite(cls)
        hist.append(tot)
        if it == 0:
            nz = [k for k in step.names if float(step.gviews[k].abs().max()) == 0.0 and k not in step.tg.no_grad_params]
            assert not nz, nz[:5]
    assert min(hist[3:]) < hist[0], hist
    assert bool(torch.isfinite(step.flatP).all())
