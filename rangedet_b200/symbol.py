"""Config-driven graph builders: the Python graph surface of the reference over the sm_100a kernels.

The reference's config (config/rangedet/rangedet_veh_wo_aug_4_18e.py:333-339) builds its graphs with

    backbone = DLABackbone(BackboneParam)
    rpn_head = RangeRpnHead(RpnParam)
    detector = RangeRCNN(RpnParam)
    train_sym = detector.get_train_symbol(backbone, rpn_head)
    test_sym  = detector.get_test_symbol(backbone, rpn_head)

The classes here take the SAME parameter classes (attributes read: rangedet/symbol/backbone/dla_backbone.py:14-15,
65-89,129-175; rangedet/symbol/head/builder.py:10-15,80-97,198-266,350-422,424-534) and return graph objects with the
reference's input / output names; instead of an mx.sym they bind to the kernels:

    step = train_sym.bind(params, batch_image, optimizer=OptimizeParam.optimizer, world_size=N, allreduce=fn)
    step.set_targets(record); losses = step.train_step(input_data, coord_s1)         # rangedet_b200.train
    out  = test_sym.bind(params, batch_image)(record)     # [rec_id, fg_cls_score, decoded_bbox, keep_inds, gt_bbox, gt_class]

What the kernels implement is the shipped vehicle / pedestrian config family (DLA stage layout, one Meta-Kernel unit,
fpn_strides (1,2,4), one class, 4+4 head convs of 128 channels, VFL + smooth-L1, wnms or NMS3D).  A config that
asks for anything else raises NotImplementedError naming the attribute -- no silent approximation.
"""
import torch

from . import dla, ops, train

GRAPH_INPUTS_TRAIN = ("input_data", "coord_s1", "rpn_cls_target_s{s}", "rpn_reg_target_s{s}", "rpn_reg_weight_s{s}",
                      "reg_normalize_weight_s{s}", "range_image_mask_s{s}", "gt_bbox_{c}_for_iou_pred", "pc_vehicle_frame_s{s}")


def _need(cond, what):
    if not cond:
        raise NotImplementedError("rangedet_b200.symbol: unsupported configuration: " + what)


def _check_normalizer(norm, what):
    """The layer kernels implement per-GPU BatchNorm with batch statistics (config:56 normalizer_factory(type="local"),
    mxnext/complicate.py:26-44): eps 1e-5 + 1e-10, momentum 0.9.  A declarative spec from rangedet_b200.shim is
    checked; a foreign callable (a real mxnext closure) cannot be inspected and is taken as the shipped local BN."""
    from .shim.mxnext_complicate import Normalizer
    if isinstance(norm, Normalizer):
        _need(norm.type in ("local", "localbn"), "%s type %r (only per-GPU 'local' BatchNorm)" % (what, norm.type))
        _need(abs(norm.eps - ops.BN_EPS) < 1e-12 and abs(norm.mom - ops.BN_MOMENTUM) < 1e-12,
              "%s eps / momentum %r / %r (kernels: %r / %r)" % (what, norm.eps, norm.mom, ops.BN_EPS, ops.BN_MOMENTUM))


class DLABackbone(object):
    """dla_backbone.py:164-175.  Validates BackboneParam against the stage layout the kernels implement."""

    def __init__(self, pBackbone):
        self.p = p = pBackbone
        _need(tuple(p.fpn_strides) == (1, 2, 4), "BackboneParam.fpn_strides must be (1, 2, 4), got %r" % (p.fpn_strides,))
        _need(dict(p.num_block) == train.NUM_BLOCK, "BackboneParam.num_block %r" % (p.num_block,))
        want_f = {'res1': 64, 'res2a': 64, 'res2': 128, 'res3a': 128, 'res3': 128, 'agg1': 64, 'agg2': 128, 'agg2a': 64, 'agg3': 64}
        _need(dict(p.num_filter) == want_f, "BackboneParam.num_filter %r" % (p.num_filter,))
        _need(bool(getattr(p, "add_data_sc", False)), "BackboneParam.add_data_sc must be True")
        units = dict(getattr(p, "meta_kernel_units", {}) or {})
        _need(set(units) <= set(train.META_UNITS), "meta_kernel_units %r (supported: %r)" % (sorted(units), train.META_UNITS))
        for name, u in units.items():   # dla_backbone.py:65-89: method picked by string, called with these kwargs
            _need(u.get("meta_func_param") == "meta_baseline_bias", "%s.meta_func_param %r" % (name, u.get("meta_func_param")))
            _need(u.get("data_channels") == 64 and u.get("coord_channels") == 3 and list(u.get("channel_list")) == [32, 64]
                  and u.get("kernel_size", 3) == 3 and u.get("stride", 1) == 1, "%s: %r" % (name, u))
        _check_normalizer(getattr(p, "normalizer", None), "BackboneParam.normalizer")
        self.use_meta = bool(units)
        self.batch_image = int(p.batch_image)
        self.range_image_shape_hw = tuple(p.range_image_shape_hw)

    def get_rpn_feature(self, data=None):
        """The reference returns the three feature symbols; here the backbone is a stage of the bound graph."""
        return [self, self, self]


class RangeRpnHead(object):
    """builder.py:80-97: reads the same RpnParam attributes."""

    def __init__(self, pRpn):
        self.p = p = pRpn
        self.fp16 = bool(p.fp16)
        self.batch_size = int(p.batch_image)
        self.class_names = tuple(p.class_names)
        self.fpn_strides = tuple(p.fpn_strides)
        self.num_classes = int(p.num_classes)
        self.num_reg_delta = int(p.num_reg_delta)
        self.cls_loss_weight = float(p.loss.cls_loss_weight)
        self.reg_loss_weight = float(p.loss.reg_loss_weight)
        self.scale_loss_shift = float(p.scale_loss_shift) if self.fp16 else 1.0
        _need(self.fpn_strides == (1, 2, 4), "RpnParam.fpn_strides %r" % (self.fpn_strides,))
        _need(self.num_classes == 1 and len(self.class_names) == 1, "one class per model (num_classes %d)" % self.num_classes)
        _need(self.num_reg_delta == 8, "RpnParam.num_reg_delta must be 8")
        h = p.head
        _need((h.cls_conv_layers, h.cls_conv_channel, h.reg_conv_layers, h.reg_conv_channel) == (4, 128, 4, 128),
              "RpnParam.head must be 4 x 128-channel cls and reg convs")
        _need(not getattr(p.loss, "l1", False), "RpnParam.loss.l1 (plain L1 regression loss)")
        _need(p.loss.iou_type in ("bev", "3d"), "RpnParam.loss.iou_type %r" % (p.loss.iou_type,))
        _check_normalizer(getattr(p, "normalizer", None), "RpnParam.normalizer")     # builder.py:212-213

    def loss_hyper(self):
        """Arguments of the fused loss head (builder.py:350-422, loss.py:4-30)."""
        l = self.p.loss
        return dict(iou_type=l.iou_type, alpha=float(l.alpha), gamma=float(l.gamma),
                    smooth_l1_scalar=float(getattr(l, "smooth_l1_scalar", 1.0)), scale_loss_shift=self.scale_loss_shift,
                    cls_loss_weight=self.cls_loss_weight, reg_loss_weight=self.reg_loss_weight)


class RangeRCNN(object):
    """builder.py:10-78."""

    def __init__(self, pDet):
        self.p = pDet
        self.fpn_strides = tuple(pDet.fpn_strides)
        self.class_names = tuple(pDet.class_names)

    def get_train_symbol(self, backbone, rpn_head):
        return TrainSymbol(self, backbone, rpn_head)

    def get_test_symbol(self, backbone, rpn_head):
        return TestSymbol(self, backbone, rpn_head)


class _Symbol(object):
    def __init__(self, det, backbone, head):
        _need(isinstance(backbone, DLABackbone) and isinstance(head, RangeRpnHead), "backbone / head must come from this module")
        _need(backbone.batch_image == head.batch_size, "BackboneParam.batch_image != RpnParam.batch_image")
        self.det, self.backbone, self.head = det, backbone, head


class TrainSymbol(_Symbol):
    """get_train_symbol (builder.py:16-52): graph inputs and loss outputs named like the reference's."""

    def list_inputs(self):
        names = []
        for n in GRAPH_INPUTS_TRAIN:
            if "{s}" in n:
                names += [n.format(s=s) for s in self.det.fpn_strides]
            elif "{c}" in n:
                names += [n.format(c=c) for c in self.det.class_names]
            else:
                names.append(n)
        return names

    def list_outputs(self):   # builder.py:374-378, 417-421
        return ["rpn_cls_loss_s%d_output" % s for s in self.det.fpn_strides] + ["rpn_reg_loss_s%d_output" % s for s in self.det.fpn_strides]

    def infer_shape(self, batch_image=None):
        B = batch_image or self.head.batch_size
        H, W = self.backbone.range_image_shape_hw
        sh = {"input_data": (B, 8, H, W), "coord_s1": (B, 3, H, W)}
        for s in self.det.fpn_strides:
            for k in ("rpn_reg_target", "rpn_reg_weight", "reg_normalize_weight"):
                sh["%s_s%d" % (k, s)] = (B, 8, H, W // s)
            sh["rpn_cls_target_s%d" % s] = (B, 1, H, W // s)
            sh["range_image_mask_s%d" % s] = (B, 1, H, W // s)
            sh["pc_vehicle_frame_s%d" % s] = (B, H * W // s, 3)
        for c in self.det.class_names:
            sh["gt_bbox_%s_for_iou_pred" % c] = (B, 200, 8 if self.head.p.loss.iou_type == "bev" else 7)
        return sh

    def bind(self, params, batch_image=None, optimizer=None, world_size=1, allreduce=None, device="cuda", lr=None,
             act_dtype=None, capture=True):
        """-> train.GraphedTrainStep on `params` (reference names).  `optimizer`: OptimizeParam.optimizer
        (type 'sgd', lr, momentum, wd, clip_gradient; tools/train.py:306-319).  rescale_grad = 1/scale_loss_shift with
        fp16 (tools/train.py:359-361); the all-reduce average is 1/world_size on top.  `act_dtype`: storage type of
        activations / operands; default = what the config asks for: torch.float16 when RpnParam.fp16 (config:35; the
        reference casts the graph to fp16, dla_backbone.py:136-137), else torch.bfloat16 (the reference would run
        fp32 there, which the tensor-core kernels do not store)."""
        B = batch_image or self.head.batch_size
        H, W = self.backbone.range_image_shape_hw
        o = optimizer
        _need(o is None or o.type == "sgd", "optimizer.type %r (only 'sgd')" % (getattr(o, "type", None),))
        step = train.GraphedTrainStep(
            params, B, H, W, lr=lr if lr is not None else (o.lr if o is not None else 0.01),
            momentum=o.momentum if o is not None else 0.9, wd=o.wd if o is not None else 1e-5,
            clip_gradient=getattr(o, "clip_gradient", None) if o is not None else 35.0,
            rescale_grad=1.0 / self.head.scale_loss_shift, device=device, use_meta=self.backbone.use_meta,
            allreduce=allreduce, world_size=world_size, loss_hyper=self.head.loss_hyper(),
            gt_name="gt_bbox_%s_for_iou_pred" % self.det.class_names[0], capture=capture,
            act_dtype=act_dtype if act_dtype is not None else (torch.float16 if self.head.fp16 else torch.bfloat16))
        return step


class TestSymbol(_Symbol):
    """get_test_symbol (builder.py:54-78) + get_fpn_prediction (:424-534): forward with moving statistics, all levels
    concatenated, sigmoid, get_sorted_foreground, Decode3DBbox; then either nothing (wnms: tools/test.py runs
    processing_cxx.wnms_4c on the host copy) or NMS3D."""

    def list_inputs(self):
        s_ = self.det.fpn_strides
        return (["input_data", "coord_s1"] + ["pc_vehicle_frame_s%d" % s for s in s_] + ["range_image_mask_s%d" % s for s in s_]
                + ["rec_id", "gt_bbox_imu", "gt_class"])

    def bind(self, params, batch_image=None, device="cuda"):
        return _TestExecutor(self, params, batch_image or self.head.batch_size, device)


class _TestExecutor(object):
    def __init__(self, sym, params, B, device):
        self.sym, self.B = sym, B
        H, W = sym.backbone.range_image_shape_hw
        _need(sym.backbone.use_meta, "the inference graph is built with the Meta-Kernel unit")
        self.fwd = dla.GraphedForward(params, B, H, W, device)
        p = sym.head.p
        c = sym.det.class_names[0]
        self.pre_n = int(p.all_proposal.rpn_pre_nms_top_n[c])
        self.post_n = int(p.all_proposal.rpn_post_nms_top_n[c])
        self.nms_thr = float(p.all_proposal.nms_thr[c])
        self.wnms = bool(getattr(p, "wnms", False))

    def __call__(self, record):
        """record: dict of CUDA tensors named like list_inputs().  Returns the reference's output list
        [rec_id, fg_cls_score (B,K), proposal (B,K,10) or (B,post_n,10), keep_inds, gt_bbox_imu, gt_class]."""
        cls_logit, bbox_delta = self.fwd(record["input_data"], record["coord_s1"])
        fg_score, prop, keep = self.predict(cls_logit, bbox_delta, record)
        return [record.get("rec_id"), fg_score, prop, keep, record.get("gt_bbox_imu"), record.get("gt_class")]

    def detections(self, record, min_score=0.5, thr_lo=0.1, thr_hi=0.5, is_3d_iou=False, hash_scale=100):
        """Forward + get_fpn_prediction + the per-frame loop body of tools/test.py:178-225 with everything resident on
        the device (wnms configs): -> list over frames of (D,8) CUDA tensors [cx,cy,cz,l,w,h,heading,score].
        Arguments as pTest.min_score[class] / pTest.nms.{thr_lo,thr_hi,is_3d_iou} (config:200-215)."""
        from . import postprocess
        _need(self.wnms, "detections() is the weighted-NMS path (RpnParam.wnms)")
        _, fg_score, boxes, _, _, _ = self(record)
        return [postprocess.frame_detections_device(fg_score[b], boxes[b], min_score, thr_lo, thr_hi, is_3d_iou, hash_scale)
                for b in range(fg_score.shape[0])]

    def predict(self, cls_logit, bbox_delta, record):
        """get_fpn_prediction (builder.py:424-534) on the head outputs: per-level (B,1,H,W_l) / (B,8,H,W_l) lists."""
        s_ = self.sym.det.fpn_strides
        B = cls_logit[0].shape[0]
        # sep_level_type(concat_all_level_per_class=True), builder.py:99-153
        cls = torch.cat([c.reshape(B, -1) for c in cls_logit], 1)
        delta = torch.cat([d.reshape(B, 8, -1).transpose(1, 2) for d in bbox_delta], 1).contiguous()
        score = torch.sigmoid(cls)
        pc = torch.cat([record["pc_vehicle_frame_s%d" % s] for s in s_], 1).contiguous()
        mask = torch.cat([record["range_image_mask_s%d" % s].reshape(B, -1) for s in s_], 1).contiguous()
        fg_score, fg_delta, fg_pc = ops.get_sorted_foreground(score, delta, pc, mask, self.pre_n)
        decoded = ops.decode_3d_bbox(fg_delta, fg_pc, is_bin=False)
        if self.wnms:
            return fg_score, decoded, torch.zeros((1,), device=decoded.device)
        keep, prop = ops.nms3d(decoded, self.nms_thr, self.post_n)
        return fg_score, prop, keep
