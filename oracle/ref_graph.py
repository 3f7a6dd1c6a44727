"""Harness: run the reference's graph code (meta_kernel.py, dla_backbone.py, head/builder.py, loss.py, mxnext/*)
eagerly through oracle/mx_eager.py and return torch tensors.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

The parameter classes restate the VALUES of config/rangedet/rangedet_veh_wo_aug_4_18e.py:31-141 with fp16 = False
(the fp32 graph; fp16 only inserts casts) and the spatial size of the test case.
"""
import torch

from . import mx_eager
from .mx_eager import S

NUM_BLOCK = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
NUM_FILTER = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128, 'agg2a': 64, 'agg3': 64}
STRIDES = (1, 2, 4)


def _configs(B, H, W, normalizer, iou_type="bev"):
    class pB:
        fp16 = False
        fpn_strides = STRIDES
        batch_image = B
        range_image_shape_hw = (H, W)
        meta_kernel_units = {'res1_unit2': dict(stride=1, meta_func_param='meta_baseline_bias', data_channels=64,
                                                coord_channels=3, channel_list=[32, 64], kernel_size=3)}
        num_block = NUM_BLOCK
        num_filter = NUM_FILTER
        add_data_sc = True
    pB.normalizer = normalizer

    class pR:
        fp16 = False
        batch_image = B
        scale_loss_shift = 128
        class_names = ('veh',)
        num_classes = 1
        fpn_strides = STRIDES
        num_reg_delta = 8
        wnms = True

        class loss:
            alpha = 1
            gamma = 2
            reg_loss_weight = 8.0
            cls_loss_weight = 10.0
            smooth_l1_scalar = 3

        class head:
            cls_conv_layers = 4
            cls_conv_channel = 128
            reg_conv_layers = 4
            reg_conv_channel = 128
    pR.normalizer = normalizer
    pR.loss.iou_type = iou_type
    return pB, pR


def _bind(P, W, extra):
    """Reference parameter names; the Meta-Kernel MLP is named after the feature width (meta_kernel.py:138)."""
    b = {k: v for k, v in P.items()}
    for k in list(P):
        if "_2656_mlp" in k:
            b[k.replace("_2656_mlp", "_%d_mlp" % W)] = P[k]
    b.update(extra)
    return b


def meta_baseline_bias(data, coord, w0, b0, w1, b1, grad_out=None):
    """MetaKernel(...).meta_baseline_bias of the reference (meta_kernel.py:166-240) -> out [, grads like
    meta_kernel_ref.meta_baseline_bias_fwd_bwd]."""
    B, C, H, W = data.shape
    data = data.detach().clone().requires_grad_(True)
    ps = [p.detach().clone().requires_grad_(True) for p in (w0, b0, w1, b1)]
    n = "unit_%d_mlp" % W
    bind = {n + "0_weight": ps[0].reshape(32, 3, 1, 1), n + "0_bias": ps[1], n + "1_weight": ps[2].reshape(C, 32, 1, 1),
            n + "1_bias": ps[3]}
    with mx_eager.reference_modules() as imp:
        mk = imp("rangedet.symbol.backbone.meta_kernel")
        with mx_eager.bound(bind):
            out = mk.MetaKernel(B, H, W, fp16=False).meta_baseline_bias(
                name="unit", data=S(data), coord_data=S(coord), data_channels=C, coord_channels=3, channel_list=[32, C],
                norm=None, conv1_filter=C, kernel_size=3).t
    if grad_out is None:
        return out.detach()
    out.backward(grad_out)
    return (out.detach(), data.grad) + tuple(p.grad for p in ps)


def backbone_head(P, data, coord, training=True, targets=None, iou_type="bev", bn_momentum=0.9):
    """DLABackbone(pBackbone).get_rpn_feature + RangeRpnHead(pRpn).get_fpn_output [+ get_fpn_loss] of the reference.
    P: parameter dict (reference names; BN moving stats are updated in place when training).  Returns dict with
    cls / reg (lists of (B,1|8,H,W_l)) and, with `targets` (graph-input names of builder.py:20-37), the loss tensors
    and the gradients MakeLoss sends into cls / reg and into every parameter."""
    B, _, H, W = data.shape
    Pg = {k: (v.detach().clone().requires_grad_(not k.endswith(("_moving_mean", "_moving_var")))) for k, v in P.items()}
    extra = {"coord_s1": coord}
    with mx_eager.reference_modules() as imp:
        norm = imp("mxnext.complicate").normalizer_factory(type="localbn", ndev=1, mom=bn_momentum)
        pB, pR = _configs(B, H, W, norm, iou_type)
        dla = imp("rangedet.symbol.backbone.dla_backbone")
        hb = imp("rangedet.symbol.head.builder")
        with mx_eager.bound(_bind(Pg, W, extra), training=training) as env:
            feats = dla.DLABackbone(pB).get_rpn_feature(S(data))
            head = hb.RangeRpnHead(pR)
            if targets is None:
                cls, reg = head.get_fpn_output(feats)
                return dict(cls=[c.t.detach() for c in cls], reg=[r.t.detach() for r in reg], feats=[f.t.detach() for f in feats],
                            moving={k: v.detach() for k, v in Pg.items() if "_moving_" in k})
            t = lambda k: S(torch.as_tensor(targets[k]))
            losses = head.get_fpn_loss(
                feats, [None] * 3, [t("rpn_reg_target_s%d" % s) for s in STRIDES], [t("rpn_reg_weight_s%d" % s) for s in STRIDES],
                [t("reg_normalize_weight_s%d" % s) for s in STRIDES], [t("range_image_mask_s%d" % s) for s in STRIDES],
                {"veh": t("gt_bbox_veh_for_iou_pred")}, [t("pc_vehicle_frame_s%d" % s) for s in STRIDES])
            cls, reg = head._cls_logit, head._bbox_delta
            for x in cls + reg:
                x.t.retain_grad()
            sum(l.t.sum() for l in losses).backward()
            used = set(env.used)
    return dict(cls=[c.t.detach() for c in cls], reg=[r.t.detach() for r in reg],
                cls_loss=[l.t.detach() for l in losses[:3]], reg_loss=[l.t.detach() for l in losses[3:]],
                d_cls=[c.t.grad for c in cls], d_reg=[r.t.grad for r in reg],
                grads={k: v.grad for k, v in Pg.items() if v.grad is not None}, used=used, moving={k: v.detach() for k, v in Pg.items() if "_moving_" in k})


def fpn_prediction(cls_logit, bbox_delta, pc_list, mask_list, pre_n, post_n=200, nms_thr=0.2):
    """RangeRpnHead.get_fpn_prediction of the reference (builder.py:424-534, wnms branch) on GIVEN head outputs
    (get_fpn_output is replaced by a function returning them): -> (fg_cls_score (B,K), decoded_bbox (B,K,10))."""
    B, _, H, W = cls_logit[0].shape
    with mx_eager.reference_modules() as imp:
        norm = imp("mxnext.complicate").normalizer_factory(type="localbn", ndev=1)
        _, pR = _configs(B, H, W, norm)

        class all_proposal:
            rpn_pre_nms_top_n = {"veh": pre_n}
            rpn_post_nms_top_n = {"veh": post_n}
            nms_thr = {"veh": 0.2}
        pR.all_proposal = all_proposal
        hb = imp("rangedet.symbol.head.builder")
        head = hb.RangeRpnHead(pR)
        head.get_fpn_output = lambda feats: ([S(c) for c in cls_logit], [S(d) for d in bbox_delta])
        with mx_eager.bound({}, training=False):
            out = head.get_fpn_prediction([None] * 3, [S(torch.as_tensor(p)) for p in pc_list],
                                          [S(torch.as_tensor(m)) for m in mask_list])
    return out[0].t.detach(), out[1].t.detach()
