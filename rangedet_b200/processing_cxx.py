"""Drop-in for the reference's pybind11 module ``processing_cxx`` (operator_cxx/src_cxx/
pybinding.cpp:6-11) for the function on the hot path:

    wnms_4c(dets, thresh, thresh_vote, _3D, hash_scale) -> (list[float] of 12*K, list[int] of K)

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
