"""Drop-in for the reference's pybind11 module ``processing_cxx`` (operator_cxx/src_cxx/
pybinding.cpp:6-11):

    wnms_4c(dets, thresh, thresh_vote, _3D, hash_scale) -> (list[float] of 12*K, list[int] of K)
    assign3D_v2(pc, bbox, bbox_center, bbox_radius, mask, is_in_nlz, max_x, min_x, max_y, min_y, max_z, min_z,
                max_dist) -> int32 (N,1)                                  (assigner.h:11-87)
    get_point_num(bbox_inds_each_pt float (N,) or (N,1)) -> float32 (N,1)  (assigner.h:89-109)

Same call signature, argument meaning and return types as point4_wnms_4c (nms.h:781-794): a
C-contiguous float32 (N,12) numpy array in, two Python lists out, empty input -> two empty lists
(nms.h:464-466).  The work runs on the current CUDA device through rd_wnms_4c; host<->device copies
are part of the call, exactly like the reference call site tools/test.py:210-218 sees it.
"""
import numpy as np
import torch

from . import ops


def wnms_4c(dets, thresh, thresh_vote, _3D=False, hash_scale=100):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.size == 0:
        return [], []
    if dets.ndim != 2 or dets.shape[1] != 12:
        raise ValueError("dets must be (N,12) float32")
    if not torch.cuda.is_available():
        raise RuntimeError("processing_cxx.wnms_4c (rangedet_b200) needs a CUDA device: no CPU fallback")
    d = torch.from_numpy(dets).cuda(non_blocking=False)
    out, keep = ops.wnms_4c_device(d, thresh, thresh_vote, bool(_3D), int(hash_scale))
    return out.reshape(-1).cpu().tolist(), keep.cpu().tolist()


def _need_cuda(what):
    if not torch.cuda.is_available():
        raise RuntimeError("processing_cxx.%s (rangedet_b200) needs a CUDA device: no CPU fallback" % what)


def assign3D_v2(pc, bbox, bbox_center, bbox_radius, mask, is_in_nlz, max_x, min_x, max_y, min_y, max_z, min_z, max_dist):
    """numpy in, numpy int32 (N,1) out, like the Eigen binding (call site rangedet/core/input.py:311-320)."""
    _need_cuda("assign3D_v2")
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    if pc.ndim != 2 or pc.shape[1] != 3:
        raise ValueError("pc must be (N,3) float32")
    out = ops.assign3d_v2_device(f(pc), f(bbox), f(bbox_center), f(bbox_radius), f(mask), f(is_in_nlz), max_x, min_x, max_y,
                                 min_y, max_z, min_z, max_dist)
    return out.cpu().numpy().reshape(-1, 1)


def get_point_num(bbox_inds_each_pt):
    """float (N,) / (N,1) in -> float32 (N,1) out (call sites input.py:433-435, util_func.py:62)."""
    _need_cuda("get_point_num")
    a = np.ascontiguousarray(bbox_inds_each_pt, dtype=np.float32).reshape(-1)
    if a.size == 0:
        return np.zeros((0, 1), np.float32)
    return ops.get_point_num_device(torch.from_numpy(a).cuda()).cpu().numpy().reshape(-1, 1)
