"""Per-frame test-time post-processing: network outputs -> final detections.

Host-side mirror of the loop body of tools/test.py:178-225 (+ the helpers bbox3d_10dim_to_11dim :56-81 and
bbox3d_12dim_to_8dim :43-53): select valid NMS3D outputs or keep everything (wnms), threshold by `min_score`,
corners (10-dim) -> 11-dim [8 corner coords, yaw, z0, height] + score, weighted NMS, -> [cx, cy, cz, l, w, h, heading,
score].  The reference does this in numpy on the host around the pybind11 `wnms_4c`; here the weighted NMS is the GPU
kernel (rangedet_b200.processing_cxx.wnms_4c, bit-exact keep indices) and the two conversions are the same numpy
expressions (float32, like the reference's arrays).
"""
import numpy as np
import torch

from . import ops, processing_cxx


def bbox3d_10dim_to_11dim(bbox3d_10dim):   # tools/test.py:56-81
    b = np.array(bbox3d_10dim, dtype=np.float32)
    xy = b[:, :8]
    yaw = np.arctan2(xy[:, 1] - xy[:, 3], xy[:, 0] - xy[:, 2])
    return np.concatenate([xy, yaw[:, None], b[:, 8:9], b[:, 9:10] - b[:, 8:9]], axis=1)


def bbox3d_12dim_to_8dim(b):               # tools/test.py:43-53
    cx = np.mean(b[:, [0, 2, 4, 6]], axis=1)
    cy = np.mean(b[:, [1, 3, 5, 7]], axis=1)
    height = b[:, 10]
    cz = b[:, 9] + height / 2
    length = np.sqrt((b[:, 2] - b[:, 0]) ** 2 + (b[:, 3] - b[:, 1]) ** 2)
    width = np.sqrt((b[:, 2] - b[:, 4]) ** 2 + (b[:, 3] - b[:, 5]) ** 2)
    return np.stack([cx, cy, cz, length, width, height, b[:, 8], b[:, 11]], axis=1)


def frame_detections(cls_score, bbox_4pts, keep_inds=None, min_score=0.5, wnms=True, thr_lo=0.1, thr_hi=0.5,
                     is_3d_iou=False, hash_scale=100):
    """One frame of tools/test.py:178-225.  cls_score (K,), bbox_4pts (K,10) [or (post_n,10) with keep_inds (post_n,)
    from NMS3D when wnms is False] -> (D,8) float32 [cx,cy,cz,l,w,h,heading,score]; D = 0 if nothing passes."""
    cls_score = np.asarray(cls_score, np.float32)
    bbox_4pts = np.asarray(bbox_4pts, np.float32)
    if not wnms:                                                   # :193-197
        keep_inds = np.asarray(keep_inds)
        bbox_4pts = bbox_4pts[keep_inds != -1]
        keep_inds = keep_inds[keep_inds != -1]
        cls_score = cls_score[keep_inds]
    fg = cls_score > min_score                                     # :200
    final_score, final_4pts = cls_score[fg], bbox_4pts[fg]
    if final_4pts.shape[0] == 0:
        return np.zeros((0, 8), np.float32)
    bbox_score = np.concatenate([bbox3d_10dim_to_11dim(final_4pts), final_score[:, None]], axis=1)   # :207-208
    if wnms:                                                       # :209-217
        flat, _ = processing_cxx.wnms_4c(bbox_score, thr_lo, thr_hi, is_3d_iou, hash_scale)
        bbox_score = np.array(flat, np.float32).reshape((-1, 12))
    if bbox_score.shape[0] == 0:
        return np.zeros((0, 8), np.float32)
    return bbox3d_12dim_to_8dim(bbox_score).astype(np.float32)


def frame_detections_device(cls_score, bbox_4pts, min_score=0.5, thr_lo=0.1, thr_hi=0.5, is_3d_iou=False, hash_scale=100):
    """Device-resident twin of frame_detections (wnms branch): CUDA tensors in -- cls_score (K,), bbox_4pts (K,10), e.g.
    one frame of symbol._TestExecutor's output -- (D,8) CUDA tensor out.  The boxes never leave the device between the
    inference graph and the weighted NMS (the reference copies them to the host, tools/test.py:178-218, because its
    wnms_4c is a CPU routine); what crosses PCIe is two counts (foreground boxes, kept boxes).  Same arithmetic as the
    host path in float32; `atan2` is torch's CUDA implementation instead of numpy's, so yaw may differ in the last bit."""
    if not (cls_score.is_cuda and bbox_4pts.is_cuda):
        raise RuntimeError("frame_detections_device needs CUDA tensors (use frame_detections for host arrays)")
    fg = cls_score > min_score                                                         # tools/test.py:200
    s, b = cls_score[fg].float(), bbox_4pts[fg].float()
    if b.shape[0] == 0:
        return torch.zeros((0, 8), device=b.device)
    yaw = torch.atan2(b[:, 1] - b[:, 3], b[:, 0] - b[:, 2])                             # :56-81
    dets = torch.cat([b[:, :8], yaw[:, None], b[:, 8:9], b[:, 9:10] - b[:, 8:9], s[:, None]], 1).contiguous()
    d, _ = ops.wnms_4c_device(dets, thr_lo, thr_hi, bool(is_3d_iou), int(hash_scale))   # :209-217
    if d.shape[0] == 0:
        return torch.zeros((0, 8), device=b.device)
    cx, cy = d[:, [0, 2, 4, 6]].mean(1), d[:, [1, 3, 5, 7]].mean(1)                     # :43-53
    length = torch.sqrt((d[:, 2] - d[:, 0]) ** 2 + (d[:, 3] - d[:, 1]) ** 2)
    width = torch.sqrt((d[:, 2] - d[:, 4]) ** 2 + (d[:, 3] - d[:, 5]) ** 2)
    return torch.stack([cx, cy, d[:, 9] + d[:, 10] / 2, length, width, d[:, 10], d[:, 8], d[:, 11]], 1)
