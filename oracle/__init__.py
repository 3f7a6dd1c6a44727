"""oracle/ -- CPU restatement of the reference's algorithms for the RangeDet hot path.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by rangedet_b200/ (the
product), which has no CPU fallback and fails loudly without its CUDA library.

Parity status
  * post-process ops (decode / rotated IoU / wNMS): restatement in rd_oracle.cpp, PINNED
    against the reference's own C++ compiled from source (oracle/_ref/librd_ref.so, built
    by build_ref.py) and against tests/golden/*.npz generated from it.
  * Meta-Kernel / convs / loss head: the arithmetic lives in MXNet (mxnet==2.0.0 per the reference's
    requirements.txt:2), absent from /root/reference and not installable here -> operator-level
    parity unpinned; meta_kernel_ref.py / dla_ref.py / dla_train_ref.py / loss_ref.py restate the graphs
    op-for-op in torch fp32 and are checked against the reference's OWN graph code executed eagerly
    (mx_eager.py stand-in for the MXNet symbol operators; ref_graph.py, ref_py.py; tests/test_reference_graph.py).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "librd_oracle.so")
_SRC = os.path.join(HERE, "rd_oracle.cpp")

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build_oracle(force=False):
    if (not force and os.path.isfile(_ORACLE_SO)
            and os.path.getmtime(_ORACLE_SO) > os.path.getmtime(_SRC)):
        return _ORACLE_SO
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                           _SRC, "-o", _ORACLE_SO])
    return _ORACLE_SO


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class _Ops:
    """Common numpy front-end over either librd_oracle.so (prefix 'orc') or librd_ref.so ('ref')."""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix
        g = lambda name: getattr(lib, prefix + "_" + name)
        self._decode = g("decode_3d_bbox")
        self._riou = g("rotated_iou")
        self._ovl = g("single_overlap")
        self._ovl.restype = ctypes.c_float
        self._wnms = g("wnms_4c")
        self._wnms.restype = ctypes.c_int

    def decode_3d_bbox(self, delta, pc, is_bin=False):
        delta, pc = _c32(delta), _c32(pc)
        box_type = delta.shape[-1]
        n = int(np.prod(delta.shape[:-1]))
        out = np.empty(delta.shape[:-1] + (10,), np.float32)
        if self.prefix == "ref":
            self._decode(_fp(delta), _fp(pc), _fp(out), ctypes.c_long(n), box_type, int(is_bin))
        else:
            self._decode(_fp(delta), _fp(pc), _fp(out), ctypes.c_long(n), int(is_bin))
        return out

    def rotated_iou(self, b1, b2, use_omp=False):
        b1, b2 = _c32(b1), _c32(b2)
        out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
        if self.prefix == "ref":
            self._riou(_fp(b1), _fp(b2), _fp(out), ctypes.c_long(b1.shape[0]),
                       ctypes.c_long(b2.shape[0]), b1.shape[1], int(use_omp))
        else:
            self._riou(_fp(b1), _fp(b2), _fp(out), ctypes.c_long(b1.shape[0]),
                       ctypes.c_long(b2.shape[0]), b1.shape[1])
        return out

    def nms3d_iou(self, box_a, box_b, normal_iou=False):
        """IoU of two 10-dim boxes as NMS3D evaluates it (nms_3d.cu:342-378)."""
        f = getattr(self.lib, self.prefix + "_nms3d_iou")
        f.restype = ctypes.c_float
        box_a, box_b = _c32(box_a), _c32(box_b)
        return float(f(_fp(box_a), _fp(box_b), int(bool(normal_iou))))

    def nms3d_kernels(self, boxes, iou_thres, max_keep, normal_iou=False):
        """(reference library only) NMS3D through the reference's own kernels emulated on the host."""
        boxes = _c32(boxes)
        B, N, _ = boxes.shape
        keep = np.empty((B, max_keep), np.int32)
        out = np.empty((B, max_keep, 10), np.float32)
        self.lib.ref_nms3d_kernels(_fp(boxes), B, N, ctypes.c_float(iou_thres), int(max_keep), int(bool(normal_iou)), _ip(keep), _fp(out))
        return keep, out

    def single_overlap(self, box1, box2, is3d=False):
        box1, box2 = _c32(box1), _c32(box2)
        return float(self._ovl(_fp(box1), _fp(box2), int(is3d)))

    def wnms_4c(self, dets, thresh, thresh_vote, is3d=False, hash_scale=100):
        dets = _c32(dets)
        n = dets.shape[0]
        out = np.empty((max(n, 1), 12), np.float32)
        keep = np.empty((max(n, 1),), np.int32)
        k = self._wnms(_fp(dets), n, ctypes.c_float(thresh), ctypes.c_float(thresh_vote), int(is3d),
                       int(hash_scale), _fp(out), _ip(keep))
        return out[:k].copy(), keep[:k].copy()


class _OracleOps(_Ops):
    def __init__(self, lib):
        super().__init__(lib, "orc")
        self._bmax = lib.orc_batch_rotated_iou_max

    def batch_rotated_iou_max(self, proposal, gt, iou_type="bev"):
        proposal, gt = _c32(proposal), _c32(gt)
        B, N, _ = proposal.shape
        out = np.empty((B, N), np.float32)
        self._bmax(_fp(proposal), _fp(gt), _fp(out), B, ctypes.c_long(N), gt.shape[1],
                   int(iou_type == "3d"))
        return out


    def assign3d_v2(self, pc, bbox, center, radius, mask, nlz, max_x, min_x, max_y, min_y, max_z, min_z, max_dist):
        pc, bbox, center = _c32(pc).reshape(-1, 3), _c32(bbox).reshape(-1, 24), _c32(center).reshape(-1, 3)
        radius, mask, nlz = _c32(radius).reshape(-1), _c32(mask).reshape(-1), _c32(nlz).reshape(-1)
        out = np.empty((pc.shape[0],), np.int32)
        cf = ctypes.c_float
        self.lib.orc_assign3d_v2(_fp(pc), _fp(bbox), _fp(center), _fp(radius), _fp(mask), _fp(nlz), cf(max_x), cf(min_x),
                                 cf(max_y), cf(min_y), cf(max_z), cf(min_z), cf(max_dist), ctypes.c_long(pc.shape[0]),
                                 int(bbox.shape[0]), _ip(out))
        return out

    def get_point_num(self, inds):
        inds = _c32(inds).reshape(-1)
        out = np.empty_like(inds)
        self.lib.orc_get_point_num(_fp(inds), ctypes.c_long(inds.shape[0]), _fp(out))
        return out

    def nms3d(self, boxes, iou_thres, max_keep, normal_iou=False):
        boxes = _c32(boxes)
        B, N, _ = boxes.shape
        keep = np.empty((B, max_keep), np.int32)
        out = np.empty((B, max_keep, 10), np.float32)
        self.lib.orc_nms3d(_fp(boxes), B, N, ctypes.c_float(iou_thres), int(max_keep), int(bool(normal_iou)), _ip(keep),
                           _fp(out))
        return keep, out


_oracle = None
_ref = None


def oracle():
    """Our restatement (always available: compiled on demand with g++)."""
    global _oracle
    if _oracle is None:
        _oracle = _OracleOps(ctypes.CDLL(build_oracle()))
    return _oracle


def reference():
    """The reference's own C++ (oracle/_ref/librd_ref.so) or None if neither the reference
    sources nor a prebuilt copy are present."""
    global _ref
    if _ref is None:
        from . import build_ref
        so = build_ref.build()
        if so is None:
            return None
        _ref = _Ops(ctypes.CDLL(so), "ref")
    return _ref
