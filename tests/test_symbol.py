"""Config-driven graph builders (rangedet_b200/symbol.py): the parameter classes below restate the VALUES of
config/rangedet/rangedet_veh_wo_aug_4_18e.py:31-141,177-184 (the file itself imports mxnet and cannot be imported
here); the builders must accept them, name inputs / outputs like rangedet/symbol/head/builder.py:16-52,374-421 and
refuse what the kernels do not implement."""
import numpy as np
import pytest
import torch

from rangedet_b200 import symbol, synth


def shipped_config(is_train=True, hw=(64, 2656)):
    class General:
        batch_image = 2 if is_train else 1           # :32
        fp16 = True                                  # :35
        scale_loss_shift = 128                       # :36
        class_names = ('veh',)                       # :41
        num_classes = 1

    class BackboneParam:                             # :89-108
        fp16 = General.fp16
        normalizer = None
        fpn_strides = (1, 2, 4)
        batch_image = General.batch_image
        range_image_shape_hw = hw
        meta_kernel_units = {'res1_unit2': dict(stride=1, meta_func_param='meta_baseline_bias', data_channels=64,
                                                coord_channels=3, channel_list=[32, 64], kernel_size=3)}
        num_block = {'res1': 2, 'res2a': 3, 'res2': 3, 'res3a': 5, 'res3': 5, 'agg1': 2, 'agg2': 2, 'agg2a': 1, 'agg3': 2}
        num_filter = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128,
                      'agg2a': 64, 'agg3': 64}
        add_data_sc = True

    class RpnParam:                                  # :110-141
        fp16 = General.fp16
        normalizer = None
        batch_image = General.batch_image
        scale_loss_shift = General.scale_loss_shift
        class_names = General.class_names
        num_classes = General.num_classes
        fpn_strides = (1, 2, 4)
        num_reg_delta = 8
        wnms = True

        class loss:
            alpha = 1
            gamma = 2
            reg_loss_weight = 8.0
            cls_loss_weight = 10.0
            iou_type = 'bev'
            smooth_l1_scalar = 3

        class head:
            cls_conv_layers = 4
            cls_conv_channel = 128
            reg_conv_layers = 4
            reg_conv_channel = 128

        class all_proposal:
            rpn_pre_nms_top_n = {'veh': 50000, 'ped': 5000, 'cyc': 5000}
            rpn_post_nms_top_n = {'veh': 200, 'ped': 200, 'cyc': 100}
            nms_thr = {'veh': 0.2, 'ped': 0.2, 'cyc': 0.2}

    class optimizer:                                 # :178-184
        type = "sgd"
        lr = 0.01 / 8 * 1 * General.batch_image * 5
        momentum = 0.9
        wd = 0.00001
        clip_gradient = 35

    return BackboneParam, RpnParam, optimizer


def _graphs(is_train=True, hw=(64, 2656)):
    pB, pR, opt = shipped_config(is_train, hw)
    backbone, head, det = symbol.DLABackbone(pB), symbol.RangeRpnHead(pR), symbol.RangeRCNN(pR)
    return det.get_train_symbol(backbone, head), det.get_test_symbol(backbone, head), opt, pB, pR


def test_shipped_config_builds_and_names_match_the_reference():
    tr, te, opt, pB, pR = _graphs()
    assert tr.list_outputs() == ["rpn_cls_loss_s1_output", "rpn_cls_loss_s2_output", "rpn_cls_loss_s4_output",
                                 "rpn_reg_loss_s1_output", "rpn_reg_loss_s2_output", "rpn_reg_loss_s4_output"]
    names = tr.list_inputs()
    for n in ("input_data", "coord_s1", "rpn_reg_target_s4", "reg_normalize_weight_s2", "range_image_mask_s1",
              "pc_vehicle_frame_s4", "gt_bbox_veh_for_iou_pred", "rpn_cls_target_s1"):
        assert n in names
    sh = tr.infer_shape()
    assert sh["input_data"] == (2, 8, 64, 2656) and sh["pc_vehicle_frame_s4"] == (2, 64 * 664, 3)
    assert sh["gt_bbox_veh_for_iou_pred"] == (2, 200, 8)
    rec = synth.rpn_targets(1, seed=0, n_vehicles=3, h=8, w=250, w_pad=256)
    assert set(rec) <= set(names)                        # the synthetic roidb record uses the graph's input names
    assert symbol.RangeRpnHead(pR).loss_hyper() == dict(iou_type='bev', alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0,
                                                        scale_loss_shift=128.0, cls_loss_weight=10.0, reg_loss_weight=8.0)
    assert "rec_id" in te.list_inputs() and "range_image_mask_s2" in te.list_inputs()


def test_unsupported_configurations_fail_loudly():
    pB, pR, _ = shipped_config()

    class B2(pB):
        num_block = dict(pB.num_block, res3=4)
    with pytest.raises(NotImplementedError, match="num_block"):
        symbol.DLABackbone(B2)

    class B3(pB):
        meta_kernel_units = {'res1_unit2': dict(pB.meta_kernel_units['res1_unit2'], meta_func_param='meta_sum')}
    with pytest.raises(NotImplementedError, match="meta_func_param"):
        symbol.DLABackbone(B3)

    class R2(pR):
        num_classes = 3
        class_names = ('veh', 'ped', 'cyc')
    with pytest.raises(NotImplementedError, match="one class"):
        symbol.RangeRpnHead(R2)

    class L1(pR.loss):
        l1 = True

    class R3(pR):
        loss = L1
    with pytest.raises(NotImplementedError, match="l1"):
        symbol.RangeRpnHead(R3)

    class R4(pR):
        fp16 = False
    assert symbol.RangeRpnHead(R4).scale_loss_shift == 1.0     # builder.py:97


@pytest.mark.gpu
def test_train_and_test_symbols_run_on_the_device():
    from rangedet_b200.model_params import make_params
    H, W = 8, 256
    tr, te, opt, pB, pR = _graphs(True, (H, W))
    P = make_params(seed=0, device="cuda")
    step = tr.bind(P, batch_image=1, optimizer=opt)
    rec = synth.rpn_targets(1, seed=3, n_vehicles=4, h=H, w=W - 6, w_pad=W)
    step.set_targets(rec)
    g = torch.Generator(device="cuda").manual_seed(0)
    data = torch.randn((1, 8, H, W), device="cuda", generator=g)
    xyz = torch.from_numpy(rec["pc_vehicle_frame_s1"]).cuda().reshape(1, H, W, 3).permute(0, 3, 1, 2).contiguous()
    losses = step.train_step(data, xyz / 25.0)
    torch.cuda.synchronize()
    assert len(losses) == 3 and all(bool(torch.isfinite(l["cls_loss"]).all()) for l in losses)
    assert float(step.hyper[0]) == pytest.approx(opt.lr) and float(step.hyper[2]) == pytest.approx(1 / 128)

    # test-time graph: outputs in the reference's order, scores descending, boxes decoded from the kept points
    class R(pR):
        class all_proposal:
            rpn_pre_nms_top_n = {'veh': 300}
            rpn_post_nms_top_n = {'veh': 50}
            nms_thr = {'veh': 0.2}
    det = symbol.RangeRCNN(R)
    ex = det.get_test_symbol(symbol.DLABackbone(pB), symbol.RangeRpnHead(R)).bind(P, batch_image=1)
    record = {k: torch.from_numpy(v).cuda() for k, v in rec.items()}
    record.update(input_data=data, coord_s1=xyz / 25.0, rec_id=torch.zeros(1), gt_bbox_imu=None, gt_class=None)
    rec_id, score, boxes, keep, _, _ = ex(record)
    assert score.shape == (1, 300) and boxes.shape == (1, 300, 10) and keep.shape == (1,)
    s = score[0].cpu().numpy()
    assert np.all(s[:-1] >= s[1:]) and 0.0 <= s.min() and s.max() <= 1.0
    assert bool(torch.isfinite(score).all())   # (boxes of a random-init head may overflow exp(): no finiteness claim)


def test_contrib_operator_names_dispatch_and_validate():
    """Operator surface by name (rangedet_b200/contrib.py): unknown CustomOp names fail like MXNet's registry, string
    kwargs are accepted, and without a device the call fails loudly instead of falling back."""
    from rangedet_b200 import contrib
    with pytest.raises(ValueError, match="not registered"):
        contrib.Custom(op_type="roi_align")
    assert contrib._as_bool("True") and not contrib._as_bool("0") and contrib._as_bool(1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            contrib.Decode3DBbox(torch.zeros(1, 4, 8), torch.zeros(1, 4, 3), is_bin="False")
        with pytest.raises(RuntimeError):
            contrib.Custom(op_type="get_sorted_foreground", cls_score=torch.zeros(1, 8), bbox_delta=torch.zeros(1, 8, 8),
                           pc=torch.zeros(1, 8, 3), mask=torch.zeros(1, 8), num_fgs="4")


@pytest.mark.gpu
def test_contrib_operators_match_ops():
    from rangedet_b200 import contrib, ops
    d, pc = synth.decode_inputs(2, 500, seed=1)
    d, pc = torch.from_numpy(d).cuda(), torch.from_numpy(pc).cuda()
    dec = contrib.Decode3DBbox(d, pc, is_bin=False)
    assert torch.equal(dec, ops.decode_3d_bbox(d, pc))
    gt = torch.from_numpy(synth.gt_boxes8(2)).cuda()
    iou = contrib.Custom(proposal=dec, gt_bbox=gt, op_type="batch_rotated_iou", iou_type="bev")
    assert torch.equal(iou, ops.batch_rotated_iou(dec, gt, "bev"))
    m = contrib.RotatedIOU(dec[0, :50, :8].contiguous(), gt[0, :20].contiguous())
    assert m.shape == (50, 20) and float(m.max()) <= 1.0
    score = torch.rand(2, 500, device="cuda")
    s, fd, fp = contrib.Custom(cls_score=score, bbox_delta=d, pc=pc, mask=torch.ones(2, 500, device="cuda"),
                               op_type="get_sorted_foreground", num_fgs="100")
    assert s.shape == (2, 100) and fd.shape == (2, 100, 8) and fp.shape == (2, 100, 3)
    keep, boxes = contrib.NMS3D(contrib.Decode3DBbox(fd, fp), 0.2, 30)
    assert keep.shape == (2, 30) and boxes.shape == (2, 30, 10) and keep.dtype == torch.int32


def _prediction_case():
    B, H, W = 2, 64, 32
    g = torch.Generator().manual_seed(0)
    cls = [torch.randn(B, 1, H, W // s, generator=g) for s in (1, 2, 4)]
    T = synth.rpn_targets(B, seed=3, n_vehicles=6, h=H, w=W - 2, w_pad=W)
    reg = [torch.from_numpy(T["rpn_reg_target_s%d" % s]) + 0.2 * torch.randn(B, 8, H, W // s, generator=g) for s in (1, 2, 4)]
    return cls, reg, T


def test_prediction_golden_is_the_reference_graph_output():
    """tests/golden/fpn_prediction.npz = RangeRpnHead.get_fpn_prediction of the reference (builder.py:424-534) executed
    through oracle/mx_eager.py on the head outputs of _prediction_case(); re-derived live where /root/reference exists."""
    from conftest import golden
    from oracle import mx_eager, ref_graph
    g = golden("fpn_prediction.npz")
    assert g["score"].shape == (2, 300) and g["boxes"].shape == (2, 300, 10)
    assert np.all(g["score"][:, :-1] >= g["score"][:, 1:])
    if not mx_eager.available():
        pytest.skip("/root/reference not present")
    cls, reg, T = _prediction_case()
    sc, box = ref_graph.fpn_prediction(cls, reg, [T["pc_vehicle_frame_s%d" % s] for s in (1, 2, 4)],
                                       [T["range_image_mask_s%d" % s].reshape(2, -1) for s in (1, 2, 4)], 300)
    assert np.array_equal(sc.numpy(), g["score"]) and np.array_equal(box.numpy(), g["boxes"])


@pytest.mark.gpu
def test_test_executor_prediction_matches_reference_graph():
    """Our inference executor's post-forward part (level concat order, sigmoid, get_sorted_foreground, decode) on the
    same head outputs: scores 1e-6 (sigmoid ulp), selected points identical, boxes 1e-5."""
    from conftest import golden
    cls, reg, T = _prediction_case()
    _, pR, _ = shipped_config(False, (64, 32))

    class R(pR):
        class all_proposal:
            rpn_pre_nms_top_n = {'veh': 300}
            rpn_post_nms_top_n = {'veh': 200}
            nms_thr = {'veh': 0.2}
    ex = symbol._TestExecutor.__new__(symbol._TestExecutor)          # the post-forward part needs no parameters
    ex.sym = symbol.RangeRCNN(R).get_test_symbol(symbol.DLABackbone(shipped_config(False, (64, 32))[0]), symbol.RangeRpnHead(R))
    ex.pre_n, ex.post_n, ex.nms_thr, ex.wnms = 300, 200, 0.2, True
    record = {k: torch.from_numpy(v).cuda() for k, v in T.items()}
    score, boxes, keep = ex.predict([c.cuda() for c in cls], [r.cuda() for r in reg], record)
    g = golden("fpn_prediction.npz")
    assert np.abs(score.cpu().numpy() - g["score"]).max() <= 1e-6
    assert np.allclose(boxes.cpu().numpy(), g["boxes"], rtol=1e-5, atol=1e-4)
    assert keep.shape == (1,)
