"""CPU restatement (numpy) of the reference's per-frame training-target generation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: **parity unpinned** -- the assigner needs Eigen
(absent) and the target arithmetic is numpy in the reference's loader with no test pinning it.  Follows
  rangedet/core/input.py:296-322   Bbox3dAssigner.get_faster_bbox3d_ind_assigner (arguments of assign3D_v2)
  rangedet/core/input.py:430-449   get_normalization_weight / get_rpn_reg_weight
  rangedet/core/input.py:452-506   get_rpn_reg_target (+ rot_alone_z :508-519)
The index work (assign3D_v2 / get_point_num) is restated in rd_oracle.cpp.
"""
import numpy as np

from . import oracle


def assigner_args(gt_corners, n_boxes_pad_radius=100.0, max_dist=20.0):
    """input.py:303-313: radius 100 per box, centre = mean of the 8 corners, GT extent, max_dist 20."""
    g = np.asarray(gt_corners, np.float32).reshape(-1, 8, 3)
    radius = (np.ones((len(g),)) * n_boxes_pad_radius).astype(np.float32)
    center = g.mean(axis=1)
    ext = [float(g[:, :, 0].max()), float(g[:, :, 0].min()), float(g[:, :, 1].max()), float(g[:, :, 1].min()),
           float(g[:, :, 2].max()), float(g[:, :, 2].min())]
    return center, radius, ext, float(max_dist)


def bbox3d_ind(pc, gt_corners, mask):
    center, radius, ext, max_dist = assigner_args(gt_corners)
    nlz = np.zeros((pc.shape[0],), np.float32)
    return oracle().assign3d_v2(pc, gt_corners, center, radius, mask, nlz, *ext, max_dist)


def normalization_weight(ind):   # input.py:430-437
    num = oracle().get_point_num(ind.astype(np.float32))
    w = 1 / num
    w[w == -1] = 0
    return w


def rpn_reg_weight(ind, reg_dim_weights):   # input.py:439-449
    out = np.zeros((ind.shape[0], len(reg_dim_weights)), np.float32)
    out[ind > -1] = np.asarray(reg_dim_weights, np.float32)
    return out


def rpn_reg_target(pc, gt_box7, ind):   # input.py:452-506 (delta_bottom_height=False)
    pc = np.asarray(pc, np.float32).reshape(-1, 3)
    inbox = ind > -1
    if inbox.sum() == 0:
        return np.zeros((pc.shape[0], 8), np.float32)
    box = np.asarray(gt_box7, np.float32)[ind]          # -1 picks the last box; zeroed below
    azimuth = np.arctan2(pc[:, 1], pc[:, 0])
    dyaw = box[:, -1] - azimuth
    c, s = np.cos(azimuth), np.sin(azimuth)
    d = box[:, :3] - pc
    rx = c * d[:, 0] + s * d[:, 1]                       # clockwise rotation about z
    ry = -s * d[:, 0] + c * d[:, 1]
    sx = np.sqrt(np.abs(rx)) * np.sign(rx)
    sy = np.sqrt(np.abs(ry)) * np.sign(ry)
    t = np.stack((sx, sy, np.log(box[:, 4]), np.log(box[:, 3]), np.cos(dyaw), np.sin(dyaw), box[:, 2] - box[:, 5] / 2,
                  np.log(box[:, 5])), axis=1).astype(np.float32)
    t[~inbox] = 0
    return t
