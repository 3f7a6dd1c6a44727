"""Regenerates tests/golden/*.npz.  Run HERE (build container, /root/reference present):

    python tests/golden/make_golden.py

Post-process vectors come from the reference's OWN C++ (oracle/_ref/librd_ref.so, compiled from
/root/reference by oracle/build_ref.py).  Meta-Kernel vectors come from the torch fp32 restatement
(oracle/meta_kernel_ref.py) -- the reference's arithmetic for that part lives in MXNet, which is not
available ("parity unpinned" at that boundary).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import meta_kernel_ref  # noqa: E402
from rangedet_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def loss_vectors():
    """RPN loss head (torch fp32 restatement, oracle/loss_ref.py; parity unpinned at the MXNet boundary)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_rpn_loss as T
    items = {}
    for it in ("bev", "3d"):
        r = T.ref_level(T.loss_case(seed=0, iou_type=it), it)
        items.update({it + "_" + k: v.numpy() for k, v in r.items()})
    np.savez_compressed(os.path.join(OUT, "rpn_loss.npz"), **items)


def assign_vectors():
    """Target assignment (restatement of assigner.h / input.py; parity unpinned: Eigen absent)."""
    from oracle import target_ref
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=12, seed=2, h=16, w=400)
    ind = target_ref.bbox3d_ind(pc, c24, mask)
    np.savez_compressed(os.path.join(OUT, "assign.npz"), ind=ind, norm_w=target_ref.normalization_weight(ind),
                        target=target_ref.rpn_reg_target(pc, b7, ind))


def nms3d_iou_vectors(ref):
    """IoU as NMS3D evaluates it, from the reference's nms_3d.cu device helpers compiled for the host."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_pinning as T
    b, pairs = T._nms3d_pairs()
    iou = np.array([[ref.nms3d_iou(b[i], b[j], nrm) for nrm in (0, 1)] for i, j in pairs], np.float32)
    np.savez_compressed(os.path.join(OUT, "nms3d_iou.npz"), iou=iou)
    # whole op through the reference's kernels emulated on the host
    b = T._nms3d_case()
    items = {}
    for tag, thr, mk, nrm in (("bev", 0.1, 300, False), ("normal", 0.3, 100, True)):
        k, bx = ref.nms3d_kernels(b, thr, mk, nrm)
        items[tag + "_keep"], items[tag + "_boxes"] = k, bx
    np.savez_compressed(os.path.join(OUT, "nms3d_keep.npz"), **items)


def ref_py_vectors():
    """Outputs of the reference's own Python CustomOps run through oracle/ref_py.py (needs /root/reference)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_pinning as T
    from oracle import ref_py
    prop, gt8, g7, score, d, pc, mask = T._ref_py_case()
    fs, fd, fp = ref_py.get_sorted_foreground(score, d, pc, mask, 200)
    np.savez_compressed(os.path.join(OUT, "ref_py_ops.npz"), bev=ref_py.batch_rotated_iou(prop, gt8, "bev"),
                        iou3d=ref_py.batch_rotated_iou(prop, g7.copy(), "3d"), fg_score=fs, fg_delta=fd, fg_pc=fp)


def prediction_vectors():
    """get_fpn_prediction of the reference executed eagerly (oracle/mx_eager.py) on fixed head outputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_symbol as T
    from oracle import ref_graph
    cls, reg, rec = T._prediction_case()
    sc, box = ref_graph.fpn_prediction(cls, reg, [rec["pc_vehicle_frame_s%d" % s] for s in (1, 2, 4)],
                                       [rec["range_image_mask_s%d" % s].reshape(2, -1) for s in (1, 2, 4)], 300)
    np.savez_compressed(os.path.join(OUT, "fpn_prediction.npz"), score=sc.numpy(), boxes=box.numpy())


def loader_vectors():
    """Bbox3dAssigner + GenerateTarget of the reference (rangedet/core/input.py) run through oracle/ref_py.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_assign as T
    from oracle import ref_py
    pc, mask, b7, c24 = T._loader_case()
    ind, tgt, nw, rw = ref_py.loader_targets(pc.reshape(64, 2650, 3), mask.reshape(64, 2650, 1), c24.reshape(-1, 8, 3), b7,
                                             T.REG_W_SHIPPED)
    sel = ind >= 0
    np.savez_compressed(os.path.join(OUT, "loader_targets.npz"), ind=ind.astype(np.int32), target_fg=tgt.reshape(-1, 8)[sel],
                        norm_fg=nw.reshape(-1, 8)[sel][:, 0], weight_fg=rw.reshape(-1, 8)[sel])


def main():
    loader_vectors()
    prediction_vectors()
    ref_py_vectors()
    loss_vectors()
    assign_vectors()
    nms3d_iou_vectors(oracle.reference())
    ref = oracle.reference()
    assert ref is not None, "needs /root/reference"
    # decode (8-dim and bin)
    d, pc = synth.decode_inputs(2, 300, seed=4)
    dbin = np.concatenate([d[..., :2], d[..., 6:7], d[..., 2:4], d[..., 7:8],
                           np.arctan2(d[..., 5:6], d[..., 4:5])], -1).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "decode.npz"), delta=d, pc=pc, out=ref.decode_3d_bbox(d, pc),
                        delta_bin=dbin, out_bin=ref.decode_3d_bbox(dbin, pc, is_bin=True))
    # rotated IoU, three box types; clustered so that many pairs overlap; plus padded GT rows
    b7 = synth.boxes7(600, seed=11, clustered=True)
    c8 = synth.boxes7_to_corners10(b7)[:, :8]
    gt8 = np.concatenate([c8[:60] + np.float32(0.07), synth.gt_boxes8(1, 10, 40, seed=5)[0]], 0)
    x5 = np.stack([b7[:, 0], b7[:, 1], b7[:, 3], b7[:, 4], b7[:, 6]], 1)
    # Types 5/7 go through sin/cos: for bit-identical boxes the reference's inclusive point-in-box
    # test sits exactly on the boundary and flips with the last ulp of sinf/cosf (glibc vs CUDA), so
    # the second set is a jittered copy (type 8 keeps exact duplicates: no trig involved there).
    b5 = x5[100:260] + np.float32([0.013, -0.007, 0.004, 0.002, 0.003])
    b7j = b7[100:260] + np.float32([0.013, -0.007, 0.02, 0.004, 0.002, 0.01, 0.003])
    np.savez_compressed(os.path.join(OUT, "rotated_iou.npz"),
                        a8=c8, b8=gt8, iou8=ref.rotated_iou(c8, gt8),
                        a5=x5[:200], b5=b5, iou5=ref.rotated_iou(x5[:200], b5),
                        a7=b7[:200], b7=b7j, iou7=ref.rotated_iou(b7[:200], b7j))
    # weighted NMS
    items = {}
    for tag, n, cl in [("clustered", 3000, True), ("uniform", 1500, False)]:
        dets = synth.wnms_dets(n, seed=5, clustered=cl)
        od, ok = ref.wnms_4c(dets, 0.1, 0.5, False, 100)
        od3, ok3 = ref.wnms_4c(dets, 0.1, 0.5, True, 100)
        items.update({tag + "_dets": dets, tag + "_out": od, tag + "_keep": ok,
                      tag + "_out3d": od3, tag + "_keep3d": ok3})
    np.savez_compressed(os.path.join(OUT, "wnms.npz"), **items)
    # Meta-Kernel fwd + bwd (torch fp32 restatement)
    torch.manual_seed(0)
    B, C, H, W = 1, 64, 5, 28
    coord = torch.from_numpy(synth.range_image_coords(B, seed=0, h=H, w=W - 3, w_pad=W))
    data = torch.from_numpy(synth.feature_map(B, C, seed=1, h=H, w=W - 3, w_pad=W))
    w0, b0, w1, b1 = [torch.from_numpy(p) for p in synth.meta_mlp_params(seed=2)]
    go = torch.randn(B, C * 9, H, W)
    out, gd, gw0, gb0, gw1, gb1 = meta_kernel_ref.meta_baseline_bias_fwd_bwd(data, coord, w0, b0, w1, b1, go)
    np.savez_compressed(os.path.join(OUT, "meta_kernel.npz"), data=data.numpy(), coord=coord.numpy(),
                        w0=w0.numpy(), b0=b0.numpy(), w1=w1.numpy(), b1=b1.numpy(), grad_out=go.numpy(),
                        out=out.numpy(), grad_data=gd.numpy(), grad_w0=gw0.numpy(), grad_b0=gb0.numpy(),
                        grad_w1=gw1.numpy(), grad_b1=gb1.numpy())
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
