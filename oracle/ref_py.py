"""Run the reference's own Python CustomOps (operator_py/*.py) in this container.

TEST INFRASTRUCTURE ONLY.  The two CustomOps on the hot path are ~40 lines of array code each on top of MXNet, which
is not installable here.  This module puts a MINIMAL stand-in for the handful of MXNet entry points they touch into
sys.modules (numpy-backed arrays; `mxnet.ndarray.contrib.RotatedIOU` = the reference's own C++ functor compiled by
oracle/build_ref.py), imports the reference modules from /root/reference UNMODIFIED, and calls their `forward`.
It therefore only works where /root/reference exists (the build container): it is used to validate the restatements
(oracle.batch_rotated_iou_max, oracle/sorted_fg_ref.py) and to generate committed golden vectors
(tests/golden/make_golden.py); nothing on the GPU box imports it.

What the stand-in assumes about MXNet (everything else is the reference's code):
  * mxnet.numpy behaves like numpy for zeros / isnan / isinf / arctan2 / concatenate / mean / sum / max / ** and
    boolean-mask assignment (mxnet.numpy's stated contract);
  * nd.topk(x, axis=1, k, ret_typ='both') returns the k largest values per row in descending order with float
    indices, nd.argsort(x, axis=0, is_ascend=False) float indices of a descending sort, both STABLE for equal keys
    (MXNet leaves the order of ties unspecified -- parity for ties stays unpinned);
  * NDArray indexing accepts float index arrays (MXNet index arrays are float32 by default).
"""
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("RD_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "operator_py", "batch_rotated_iou.py"))


class Arr(np.ndarray):
    """numpy array with the few NDArray methods the ops call."""
    context = "cpu(0)"

    def as_nd_ndarray(self):
        return self

    def as_np_ndarray(self):
        return self

    def __getitem__(self, idx):
        if isinstance(idx, np.ndarray) and idx.dtype.kind == "f":   # MXNet index arrays are float32
            idx = np.asarray(idx).astype(np.int64)
        return super().__getitem__(idx)


def A(x):
    return np.ascontiguousarray(x).view(Arr)


def _stub_modules(rotated_iou):
    mx = types.ModuleType("mxnet")
    mnp = types.ModuleType("mxnet.numpy")
    for name in ("isnan", "isinf", "arctan2"):
        setattr(mnp, name, getattr(np, name))
    mnp.concatenate = lambda arrays, axis=0: A(np.concatenate([np.asarray(a) for a in arrays], axis=axis))
    mnp.zeros = lambda shape, dtype=np.float32, ctx=None: A(np.zeros(shape, dtype))
    nd = types.ModuleType("mxnet.ndarray")
    nd.zeros = lambda shape, ctx=None, dtype=np.float32: A(np.zeros(shape, dtype))

    def topk(x, axis=1, k=1, ret_typ="indices"):
        assert axis == 1 and ret_typ == "both"
        order = np.argsort(-np.asarray(x), axis=1, kind="stable")[:, :k]
        return A(np.take_along_axis(np.asarray(x), order, 1)), A(order.astype(np.float32))

    def argsort(x, axis=0, is_ascend=True):
        v = np.asarray(x)
        return A(np.argsort(v if is_ascend else -v, axis=axis, kind="stable").astype(np.float32))

    nd.topk, nd.argsort = topk, argsort
    contrib = types.ModuleType("mxnet.ndarray.contrib")
    contrib.RotatedIOU = lambda a, b: A(rotated_iou(np.ascontiguousarray(a), np.ascontiguousarray(b)))
    nd.contrib = contrib
    op = types.ModuleType("mxnet.operator")

    class CustomOp(object):
        def assign(self, dst, req, src):
            dst[...] = src

    class CustomOpProp(object):
        def __init__(self, need_top_grad=False):
            self.need_top_grad = need_top_grad

    op.CustomOp, op.CustomOpProp = CustomOp, CustomOpProp
    op.register = lambda name: (lambda cls: cls)
    mx.operator, mx.nd, mx.ndarray, mx.numpy = op, nd, nd, mnp
    return {"mxnet": mx, "mxnet.numpy": mnp, "mxnet.ndarray": nd, "mxnet.ndarray.contrib": contrib, "mxnet.operator": op}


@contextlib.contextmanager
def _reference_module(fname):
    from . import reference
    ref = reference()
    assert ref is not None and available(), "needs /root/reference"
    stubs = _stub_modules(ref.rotated_iou)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_ref_" + fname[:-3], os.path.join(REF, "operator_py", fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        yield mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def batch_rotated_iou(proposal, gt_bbox, iou_type="bev"):
    """BatchRotatedIOU(iou_type).forward on numpy inputs -> (B,N) IoU target (operator_py/batch_rotated_iou.py:11-25)."""
    with _reference_module("batch_rotated_iou.py") as m:
        op = m.BatchRotatedIOU(iou_type)
        out = A(np.zeros(proposal.shape[:2], np.float32))
        op.forward(True, ["write"], [A(np.array(proposal, np.float32)), A(np.array(gt_bbox, np.float32))], [out], [])
        return np.asarray(out)


def get_sorted_foreground(cls_score, bbox_delta, pc, mask, num_fgs):
    """GetSortedFGOperator(num_fgs).forward on numpy inputs (operator_py/get_sorted_foreground.py:11-40)."""
    with _reference_module("get_sorted_foreground.py") as m:
        op = m.GetSortedFGOperator(int(num_fgs))
        B = cls_score.shape[0]
        outs = [A(np.zeros((B, num_fgs), np.float32)), A(np.zeros((B, num_fgs, 8), np.float32)), A(np.zeros((B, num_fgs, 3), np.float32))]
        op.forward(False, ["write"] * 3, [A(np.array(x, np.float32)) for x in (cls_score, bbox_delta, pc, mask)], outs, [])
        return tuple(np.asarray(o) for o in outs)


# ---- the reference's numpy data-loader transforms (rangedet/core/input.py) ---------------------------------------
@contextlib.contextmanager
def _loader_modules():
    """rangedet.core.input imports `processing_cxx` (pybind11 + Eigen: not buildable here -> the C++ restatement in
    rd_oracle.cpp stands in, so the index assignment itself stays unpinned) and utils.detection_input (needs
    mx.io.DataIter as a base class only)."""
    from . import oracle
    assert os.path.isfile(os.path.join(REF, "rangedet", "core", "input.py")), "needs /root/reference"
    orc = oracle()
    pcx = types.ModuleType("processing_cxx")
    pcx.assign3D_v2 = lambda pc, bbox, ctr, rad, mask, nlz, *f: orc.assign3d_v2(pc, bbox, ctr, rad, mask, nlz, *f).reshape(-1, 1)
    pcx.get_point_num = lambda inds: orc.get_point_num(inds).reshape(-1, 1)
    mx = types.ModuleType("mxnet")
    mx.io = types.SimpleNamespace(DataIter=object, DataBatch=object, DataDesc=object)
    stubs = {"processing_cxx": pcx, "mxnet": mx}
    pk = ("rangedet", "utils")
    saved = {k: sys.modules.get(k) for k in stubs}
    for k in list(sys.modules):
        if k.split(".")[0] in pk:
            saved[k] = sys.modules.pop(k)
    sys.modules.update(stubs)
    sys.path.insert(0, REF)
    try:
        import importlib
        yield importlib.import_module("rangedet.core.input")
    finally:
        sys.path.remove(REF)
        for k in list(sys.modules):
            if k.split(".")[0] in pk or k in stubs:
                sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def loader_targets(pc_hw3, mask_hw1, gt_corners_m83, gt_box7, reg_weight):
    """Bbox3dAssigner.apply + GenerateTarget.apply of the reference (input.py:276-372) on one 64x2650 frame ->
    (bbox3d_ind (H*W,), rpn_reg_target, reg_normalize_weight, rpn_reg_weight each (H,W,8))."""
    H, W = pc_hw3.shape[:2]
    with _loader_modules() as inp:
        class GP:
            feat_size = (H, W)
            num_classes = 1
            label_set = [1]
        GP.reg_weight = list(reg_weight)
        rec = {"pc_vehicle_frame": np.array(pc_hw3, np.float32), "gt_bbox_imu": np.array(gt_corners_m83, np.float32),
               "range_image_mask": np.array(mask_hw1), "gt_bbox_csa": np.array(gt_box7, np.float32),
               "gt_class": np.ones((len(gt_box7),), np.int32)}
        inp.Bbox3dAssigner(GP).apply(rec)
        inp.GenerateTarget(GP).apply(rec)
        return (rec["bbox3d_ind_of_each_pt"].reshape(-1), rec["rpn_reg_target"], rec["reg_normalize_weight"], rec["rpn_reg_weight"])


def test_py_functions():
    """The two numpy helpers of tools/test.py (bbox3d_12dim_to_8dim :43-53, bbox3d_10dim_to_11dim :56-81), their
    source taken verbatim from the file with `ast` (the module itself cannot be imported: it dlopens contrib_cxx.so and
    imports MXNet on line 1) and executed with numpy."""
    import ast
    path = os.path.join(REF, "tools", "test.py")
    tree = ast.parse(open(path).read())
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("bbox3d_12dim_to_8dim", "bbox3d_10dim_to_11dim"):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns["bbox3d_10dim_to_11dim"], ns["bbox3d_12dim_to_8dim"]
