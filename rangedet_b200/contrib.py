"""The operator names the reference's graph builders call, bound to the sm_100a kernels.

  mx.sym.contrib.Decode3DBbox(bbox_deltas, pc_laser_frame, is_bin=False)     operator_cxx/contrib/decode_3d_bbox.cc:15-65
  mx.nd.contrib.RotatedIOU(boxes1, boxes2)                                    operator_cxx/contrib/rotated_iou.cc:12-60
  mx.sym.contrib.NMS3D(boxes, iou_thres, max_keep, normal_iou=False)          operator_cxx/contrib/nms_3d.cc:22-68
  mx.sym.Custom(op_type='batch_rotated_iou', proposal=, gt_bbox=, iou_type=)  operator_py/batch_rotated_iou.py:71-110
  mx.sym.Custom(op_type='get_sorted_foreground', cls_score=, bbox_delta=, pc=, mask=, num_fgs=)
                                                                              operator_py/get_sorted_foreground.py:47-84
Same argument names (positional or keyword) and the same outputs; `Custom` keyword arguments may arrive as strings,
as MXNet passes CustomOp kwargs (get_sorted_foreground.py:51), and an unknown op_type raises like MXNet's registry.
All results are detached: every one of these ops has a zero gradient in the reference (MakeZeroGradNodes /
assign(in_grad, 0)).
"""
from . import ops


def Decode3DBbox(bbox_deltas, pc_laser_frame, is_bin=False, name=None):
    return ops.decode_3d_bbox(bbox_deltas, pc_laser_frame, is_bin=_as_bool(is_bin))


def RotatedIOU(boxes1, boxes2, name=None):
    return ops.rotated_iou(boxes1, boxes2)


def NMS3D(boxes, iou_thres, max_keep, normal_iou=False, name=None):
    """-> (keep_idx (B,max_keep) int32, boxes_after_nms (B,max_keep,10)), the op's two outputs in its order."""
    return ops.nms3d(boxes, float(iou_thres), int(max_keep), _as_bool(normal_iou))


def _as_bool(v):
    if isinstance(v, str):
        return v.strip().lower() in ("1", "true")
    return bool(v)


def _batch_rotated_iou(proposal, gt_bbox, iou_type="bev", name=None):
    return ops.batch_rotated_iou(proposal, gt_bbox, str(iou_type))


def _get_sorted_foreground(cls_score, bbox_delta, pc, mask, num_fgs, name=None):
    return ops.get_sorted_foreground(cls_score, bbox_delta, pc, mask, int(num_fgs))


CUSTOM_OPS = {"batch_rotated_iou": _batch_rotated_iou, "get_sorted_foreground": _get_sorted_foreground}


def Custom(*args, op_type=None, **kwargs):
    """mx.sym.Custom(..., op_type=<registered name>, **kwargs) for the two CustomOps on the hot path."""
    if op_type not in CUSTOM_OPS:
        raise ValueError("Custom operator %r is not registered (available: %s)" % (op_type, ", ".join(sorted(CUSTOM_OPS))))
    return CUSTOM_OPS[op_type](*args, **kwargs)
