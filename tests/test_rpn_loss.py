"""RPN loss head: rd_rpn_loss vs the torch-fp32 restatement of rangedet/symbol/head/
loss.py:4-30 and builder.py:155-197,350-422 (oracle/loss_ref.py).

Tolerances (north_star: 1e-3 rel for fp32 activations / regressions): loss tensors and gradients
|d| <= 1e-3*|ref| + 1e-3*rms(ref); IoU target |d| <= 1e-3 vs the CPU restatement and BIT-EXACT vs the
library's own decode -> batch_rotated_iou pair (which is pinned to the reference's compiled C++).
"""
import numpy as np
import pytest
import torch

from conftest import golden
from rangedet_b200 import synth

HYP = dict(alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0, cls_loss_weight=10.0, reg_loss_weight=8.0)


def loss_case(B=2, H=8, W=96, seed=0, iou_type="bev"):
    """One pyramid level: targets from the synthetic roidb generator, head outputs = targets + noise on the
    foreground (so IoU targets spread over (0,1)) and small random values elsewhere."""
    rng = np.random.default_rng(seed)
    T = synth.rpn_targets(B, seed=seed + 5, n_vehicles=4, strides=(1,), h=H, w=W - 3, w_pad=W, n_gt=200)
    tgt = T["rpn_reg_target_s1"]
    fg = T["rpn_reg_weight_s1"] > 0
    delta = np.where(fg, tgt + rng.normal(0, 0.15, tgt.shape), rng.normal(0, 0.5, tgt.shape)).astype(np.float32)
    logit = rng.normal(0, 2.0, (B, 1, H, W)).astype(np.float32)
    logit.reshape(-1)[:6] = [0.0, -0.0, 20.0, -20.0, 14.5, -14.5]   # clip / softrelu corners (loss.py:4-20)
    gt = T["gt_bbox_veh_for_iou_pred"]
    if iou_type == "3d":   # (B,G,7) [x,y,z,l,w,h,yaw] GT (operator_py/batch_rotated_iou.py:31-36), padded like :264-265
        g7 = np.zeros((B, 200, 7), np.float32)
        g7[:, :, 3:6] = 1e-3
        for b in range(B):
            c = gt[b, :4].reshape(4, 4, 2)
            ctr = c.mean(1)
            l = np.linalg.norm(c[:, 0] - c[:, 1], axis=1)
            w = np.linalg.norm(c[:, 1] - c[:, 2], axis=1)
            yaw = np.arctan2(c[:, 0, 1] - c[:, 1, 1], c[:, 0, 0] - c[:, 1, 0])
            g7[b, :4] = np.stack([ctr[:, 0], ctr[:, 1], np.full(4, 0.5), l, w, np.full(4, 1.7), yaw], 1)
        gt = g7
    return dict(cls_logit=logit, reg_delta=delta, pc=T["pc_vehicle_frame_s1"], gt=gt, mask=T["range_image_mask_s1"],
                reg_target=tgt, reg_weight=T["rpn_reg_weight_s1"], reg_norm_weight=T["reg_normalize_weight_s1"])


def ref_level(c, iou_type="bev", **kw):
    from oracle import loss_ref
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in c.items()}
    return loss_ref.rpn_loss_level(t["cls_logit"], t["reg_delta"], c["pc"], c["gt"], t["mask"], t["reg_target"], t["reg_weight"],
                                   t["reg_norm_weight"], iou_type=iou_type, **HYP, **kw)


# ---- CPU: the restatement against closed forms and the committed golden vector ----------------------------
def test_vfl_restatement_matches_closed_form():
    from oracle import loss_ref
    x = torch.linspace(-12, 12, 97, dtype=torch.float64)
    for t in (0.0, 0.3, 1.0):
        s = torch.full_like(x, t)
        got = loss_ref.vari_focal_loss(x, s, 1.0, alpha=1.0, gamma=2.0)
        p = torch.sigmoid(x)
        bce = -(s * torch.log(p) + (1 - s) * torch.log1p(-p))      # loss_init = 2 * (0.5 t log p + 0.5 (1-t) log(1-p))
        want = bce * s if t > 0 else bce * p ** 2
        assert torch.allclose(got, want, rtol=1e-9, atol=1e-12)


def test_smooth_l1_known_answers():
    from oracle import loss_ref
    x = torch.tensor([0.0, 0.05, -0.05, 1.0 / 9, 0.2, -2.0])
    got = loss_ref.smooth_l1(x, 3.0)
    want = torch.tensor([0.0, 0.5 * 9 * 0.0025, 0.5 * 9 * 0.0025, 0.5 * 9 / 81, 0.2 - 0.5 / 9, 2.0 - 0.5 / 9])
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-8)


def test_loss_gradients_match_finite_differences():
    c = loss_case(B=1, H=4, W=32, seed=3)
    r = ref_level(c)
    from oracle import loss_ref
    x = torch.from_numpy(c["cls_logit"]).double()
    t = r["iou_target"].double()
    m = torch.from_numpy(c["mask"]).double()
    f = lambda v: (loss_ref.vari_focal_loss(v, t, 1.0) * m / (m.sum() + 1)).sum() * 1280.0
    eps = 1e-6
    for i in (7, 19, 40, 77):
        e = torch.zeros_like(x)
        e.view(-1)[i] = eps
        fd = float((f(x + e) - f(x - e)) / (2 * eps))
        assert abs(fd - float(r["d_cls"].view(-1)[i])) <= 1e-3 * abs(fd) + 1e-5


def test_loss_restatement_matches_golden():
    g = golden("rpn_loss.npz")
    for it in ("bev", "3d"):
        c = loss_case(seed=0, iou_type=it)
        r = ref_level(c, it)
        for k in ("iou_target", "cls_loss", "reg_loss", "d_cls", "d_reg"):
            want = g[it + "_" + k]
            assert np.allclose(r[k].numpy(), want, rtol=1e-5, atol=1e-7 * max(1.0, float(np.abs(want).max()))), (it, k)
        assert float((r["iou_target"] > 0.3).float().mean()) > 0.01   # the case does exercise positives


# ---- GPU parity -------------------------------------------------------------------------------------------
def _close(a, b, what):
    a, b = a.double().cpu().numpy(), b.double().cpu().numpy()
    tol = 1e-3 * np.abs(b) + 1e-3 * np.sqrt((b ** 2).mean()) + 1e-12
    bad = np.abs(a - b) > tol
    assert not bad.any(), "%s: %d / %d outside tolerance, worst %.3e" % (what, bad.sum(), bad.size, np.abs(a - b).max())


@pytest.mark.gpu
@pytest.mark.parametrize("iou_type", ["bev", "3d"])
@pytest.mark.parametrize("shape", [(2, 8, 96), (1, 5, 131)])
def test_rpn_loss_matches_restatement(iou_type, shape):
    from rangedet_b200 import ops
    B, H, W = shape
    c = loss_case(B, H, W, seed=0 if shape[2] == 96 else 2, iou_type=iou_type)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items()}
    o = ops.rpn_loss(t["cls_logit"], t["reg_delta"], t["pc"], t["gt"], t["mask"], t["reg_target"], t["reg_weight"],
                     t["reg_norm_weight"], iou_type=iou_type, **HYP)
    r0 = ref_level(c, iou_type)
    assert float((o["iou_target"].cpu() - r0["iou_target"]).abs().max()) <= 1e-3
    # loss / gradients on the GPU's own IoU target (the reference's loss is discontinuous at target == 0)
    r = ref_level(c, iou_type, iou_target_override=o["iou_target"].cpu())
    for k in ("cls_loss", "reg_loss", "d_cls", "d_reg"):
        _close(o[k].cpu(), r[k], k)
    # fused IoU target == the library's decode -> batch_rotated_iou (bit-exact)
    d = t["reg_delta"].reshape(B, 8, -1).transpose(1, 2).contiguous()
    gt = t["gt"]
    if iou_type == "3d":
        gt = gt.clone()
    it2 = ops.batch_rotated_iou(ops.decode_3d_bbox(d, t["pc"]), gt, iou_type)
    assert torch.equal(it2.reshape(o["iou_target"].shape), o["iou_target"])


@pytest.mark.gpu
def test_rpn_loss_full_level_properties():
    """Full level-0 size (B=2, 64x2656): fused IoU target bit-identical to decode -> batch IoU max, gradients
    vanish exactly where the mask / weights are zero, gradient sums match the oracle on a row sample."""
    from rangedet_b200 import ops
    B = 2
    T = synth.rpn_targets(B, seed=11)
    g = torch.Generator().manual_seed(0)
    tgt = torch.from_numpy(T["rpn_reg_target_s1"])
    fg = torch.from_numpy(T["rpn_reg_weight_s1"]) > 0
    delta = torch.where(fg, tgt + 0.15 * torch.randn(tgt.shape, generator=g), 0.5 * torch.randn(tgt.shape, generator=g)).cuda()
    logit = (2.0 * torch.randn((B, 1, 64, 2656), generator=g)).cuda()
    c = {k: torch.from_numpy(v).cuda() for k, v in T.items()}
    o = ops.rpn_loss(logit, delta, c["pc_vehicle_frame_s1"], c["gt_bbox_veh_for_iou_pred"], c["range_image_mask_s1"],
                     c["rpn_reg_target_s1"], c["rpn_reg_weight_s1"], c["reg_normalize_weight_s1"], **HYP)
    d = delta.reshape(B, 8, -1).transpose(1, 2).contiguous()
    it2 = ops.batch_rotated_iou(ops.decode_3d_bbox(d, c["pc_vehicle_frame_s1"]), c["gt_bbox_veh_for_iou_pred"])
    assert torch.equal(it2.reshape(B, 1, 64, 2656), o["iou_target"])
    assert float(o["iou_target"].max()) > 0.5 and float(o["iou_target"].min()) >= 0.0
    assert bool((o["d_cls"][c["range_image_mask_s1"] == 0] == 0).all())
    assert bool((o["d_reg"][c["rpn_reg_weight_s1"] == 0] == 0).all())
    for k in o:
        assert bool(torch.isfinite(o[k]).all()), k
    # a horizontal band against the restatement with the same normalisers
    from oracle import loss_ref
    t = o["iou_target"].cpu()
    x = logit.cpu().requires_grad_(True)
    m = torch.from_numpy(T["range_image_mask_s1"])
    cls = loss_ref.vari_focal_loss(x, t, 1.0) * m / (m.sum() + 1)
    cls.sum().backward()
    _close(o["cls_loss"].cpu(), cls.detach(), "cls_loss")
    _close(o["d_cls"].cpu(), x.grad * 1280.0, "d_cls")


@pytest.mark.gpu
def test_rpn_loss_edge_cases():
    from rangedet_b200 import ops
    c = loss_case(1, 4, 32, seed=1)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items()}
    z = torch.zeros_like
    o = ops.rpn_loss(t["cls_logit"], t["reg_delta"], t["pc"], t["gt"], z(t["mask"]), t["reg_target"], z(t["reg_weight"]),
                     z(t["reg_norm_weight"]), **HYP)   # empty frame: normalisers = 0 + 1, all losses zero
    for k in ("cls_loss", "reg_loss", "d_cls", "d_reg"):
        assert float(o[k].abs().max()) == 0.0
    with pytest.raises(ValueError):
        ops.rpn_loss(t["cls_logit"], t["reg_delta"], t["pc"], t["gt"], t["mask"], t["reg_target"], t["reg_weight"],
                     t["reg_norm_weight"], iou_type="giou")
    with pytest.raises(ValueError):
        ops.rpn_loss(t["cls_logit"], t["reg_delta"][:, :7], t["pc"], t["gt"], t["mask"], t["reg_target"], t["reg_weight"],
                     t["reg_norm_weight"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("iou_type", ["bev", "3d"])
@pytest.mark.parametrize("shape", [(2, 8, 96), (1, 5, 131)])
def test_rpn_loss_on_the_nhwc_head_tensors_is_the_same_loss(iou_type, shape, dtype):
    """rd_rpn_loss_nhwc_* (head outputs and gradients in the NHWC 16-bit layout of the head convolutions) == rd_rpn_loss on the
    widened planar copies: loss tensors and IoU target bit-identical, gradients = the planar gradients rounded to the storage
    type, written to channel 0 / channels 0..7 only (the other channels, holding a sentinel, and the halo stay untouched)."""
    from rangedet_b200 import ops
    B, H, W = shape
    c = loss_case(B, H, W, seed=0 if shape[2] == 96 else 2, iou_type=iou_type)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items()}
    g = torch.Generator(device="cuda").manual_seed(9)
    cls_pad = torch.zeros((B, H + 2, W + 2, 64), device="cuda", dtype=dtype)
    reg_pad = torch.zeros_like(cls_pad)
    cls_pad[:, 1:-1, 1:-1] = torch.randn((B, H, W, 64), device="cuda", generator=g).to(dtype)       # channels > 0: junk the loss must ignore
    reg_pad[:, 1:-1, 1:-1] = torch.randn((B, H, W, 64), device="cuda", generator=g).to(dtype)
    cls_pad[:, 1:-1, 1:-1, 0] = t["cls_logit"][:, 0].to(dtype)
    reg_pad[:, 1:-1, 1:-1, :8] = t["reg_delta"].permute(0, 2, 3, 1).to(dtype)
    dcls, dreg = torch.full_like(cls_pad, 7.0), torch.full_like(reg_pad, 7.0)
    o = ops.rpn_loss_nhwc(cls_pad, reg_pad, t["pc"], t["gt"], t["mask"], t["reg_target"], t["reg_weight"], t["reg_norm_weight"],
                          dcls, dreg, iou_type=iou_type, **HYP)
    want = ops.rpn_loss(ops.nhwc_to_nchw(cls_pad, 1), ops.nhwc_to_nchw(reg_pad, 8), t["pc"], t["gt"], t["mask"], t["reg_target"],
                        t["reg_weight"], t["reg_norm_weight"], iou_type=iou_type, **HYP)
    for k in ("iou_target", "cls_loss", "reg_loss"):
        assert torch.equal(o[k], want[k]), k
    assert torch.equal(dcls[:, 1:-1, 1:-1, 0], want["d_cls"][:, 0].to(dtype))
    assert torch.equal(dreg[:, 1:-1, 1:-1, :8], want["d_reg"].permute(0, 2, 3, 1).to(dtype))
    assert bool((dcls[..., 1:] == 7.0).all()) and bool((dreg[..., 8:] == 7.0).all())
    assert bool((dcls[:, 0] == 7.0).all()) and bool((dreg[:, :, 0] == 7.0).all())                # halo untouched
