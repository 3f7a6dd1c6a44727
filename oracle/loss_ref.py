"""CPU restatement (torch fp32, op for op, gradients by torch autograd) of the reference's RPN loss head.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: **parity unpinned** at the MXNet boundary
(the element-wise arithmetic lives in mxnet==2.0.0, absent here; no reference test pins it); the IoU target
goes through the decode / rotated-IoU restatement that IS pinned against the reference's compiled C++.

Follows
  rangedet/symbol/head/loss.py:4-22     sigmoid_bce_loss_with_logits
  rangedet/symbol/head/loss.py:22-30    vari_focal_loss
  rangedet/symbol/head/builder.py:155-197  get_iou_target
  rangedet/symbol/head/builder.py:350-379  get_vfl_loss
  rangedet/symbol/head/builder.py:381-422  get_normalize_reg_loss
MXNet semantics restated: `MakeLoss(grad_scale=s)` forwards its input and back-propagates the constant s
per element; `clip` passes the gradient where a_min <= x <= a_max; `smooth_l1(x, scalar=s)` =
0.5 (s x)^2 if |x| < 1/s^2 else |x| - 0.5/s^2; comparison ops have zero gradient; `softrelu` = log(1+exp).
"""
import numpy as np
import torch

from . import oracle


def sigmoid_bce_loss_with_logits(logits, labels, alpha, loss_scale=1.0):  # loss.py:4-20
    p = torch.sigmoid(logits)
    ge = (logits >= 0).to(logits.dtype)
    minus_logits_mask = -1.0 * logits * ge
    negative_abs_logits = logits - 2 * logits * ge
    log_one_exp_minus_abs = torch.nn.functional.softplus(negative_abs_logits, threshold=1e9)
    minus_log = minus_logits_mask - log_one_exp_minus_abs
    alpha_labels = alpha * labels
    log_p_clip = torch.log(torch.clamp(p, 1e-6, 1 - 1e-6))
    one_alpha_one_labels = (1.0 - alpha) * (1 - labels)
    return -1 * loss_scale * (alpha_labels * log_p_clip + one_alpha_one_labels * minus_log)


def vari_focal_loss(pred, score, loss_scale, alpha=1.0, gamma=2.0):  # loss.py:22-30
    pred_sigmoid = torch.sigmoid(pred)
    loss_init = sigmoid_bce_loss_with_logits(pred, score, alpha=0.5, loss_scale=loss_scale) * 2.0
    positive_mask = (score > 0).to(pred.dtype)
    loss_positive = loss_init * score * positive_mask
    negative_mask = (score == 0).to(pred.dtype)
    loss_negative = loss_init * alpha * torch.pow(torch.abs(score - pred_sigmoid), gamma) * negative_mask
    return loss_negative + loss_positive


def smooth_l1(x, scalar):  # mx.sym.smooth_l1
    s2 = scalar * scalar
    return torch.where(x.abs() > 1.0 / s2, x.abs() - 0.5 / s2, 0.5 * x * x * s2)


def iou_target(reg_delta, pc, gt, iou_type="bev"):
    """builder.py:155-197 for one level and one class.  reg_delta (B,8,H,W) -> (B,1,H,W) (no gradient)."""
    B, _, H, W = reg_delta.shape
    bbox_delta = reg_delta.detach().reshape(B, 8, H * W).permute(0, 2, 1).contiguous().numpy()   # :129-141
    decoded = oracle().decode_3d_bbox(bbox_delta, np.asarray(pc, np.float32), is_bin=False)      # :176-178
    iou = oracle().batch_rotated_iou_max(decoded, np.asarray(gt, np.float32), iou_type)           # :179-185
    return torch.from_numpy(iou).reshape(B, 1, H, W)


def rpn_loss_level(cls_logit, reg_delta, pc, gt, mask, reg_target, reg_weight, reg_norm_weight, iou_type="bev",
                   alpha=1.0, gamma=2.0, smooth_l1_scalar=3.0, scale_loss_shift=128.0, cls_loss_weight=10.0,
                   reg_loss_weight=8.0, iou_target_override=None):
    """One pyramid level of get_fpn_loss (builder.py:300-348).  Returns a dict with iou_target, cls_loss,
    reg_loss (the graph outputs) and d_cls, d_reg = what MakeLoss back-propagates into the head outputs."""
    x = cls_logit.detach().clone().requires_grad_(True)
    d = reg_delta.detach().clone().requires_grad_(True)
    t = iou_target(d, pc, gt, iou_type) if iou_target_override is None else iou_target_override.detach().reshape(x.shape)
    # get_vfl_loss :350-379
    vfl = vari_focal_loss(x, t, 1.0, alpha=alpha, gamma=gamma)
    norm = mask.sum() + 1
    cls_loss = vfl * mask / norm
    # get_normalize_reg_loss :381-422
    reg = smooth_l1(d - reg_target, smooth_l1_scalar)
    rnorm = reg_norm_weight.sum() + 1
    reg_loss = reg * reg_weight * reg_norm_weight / rnorm * reg_loss_weight
    # MakeLoss backward: grad_scale per element
    (cls_loss.sum() * (scale_loss_shift * cls_loss_weight) + reg_loss.sum() * scale_loss_shift).backward()
    return dict(iou_target=t, cls_loss=cls_loss.detach(), reg_loss=reg_loss.detach(), d_cls=x.grad, d_reg=d.grad)
