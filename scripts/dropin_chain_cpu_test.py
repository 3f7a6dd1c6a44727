"""Test-time drop-in chain on the CPU (build container, needs /root/reference; ~1 min):

  the reference's UNMODIFIED config (is_train=False) -> its own loader transform list on a synthetic raw frame ->
    (a) the reference's inference graph (DLABackbone + get_fpn_output with moving statistics + get_fpn_prediction:
        sigmoid, get_sorted_foreground CustomOp, Decode3DBbox) executed eagerly through oracle/mx_eager.py
    (b) this package's forward (rangedet_b200.dla) + symbol._TestExecutor.predict over the emulated kernel API
  in float64 (logic only): selected foreground scores and decoded boxes side by side (ranks can swap where two scores
  differ by less than the fp32 rounding of the folded BatchNorm coefficients, so boxes are also compared as sets).
  TEST INFRASTRUCTURE.
"""
import importlib
import json
import os
import pathlib
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fake_ops  # noqa: E402
import test_reference_config as trc  # noqa: E402
from oracle import dla_ref, oracle, ref_graph  # noqa: E402
from rangedet_b200 import dla, symbol  # noqa: E402


def main():
    orc = oracle()
    fake_ops.set_exact(True)
    dla.ops = symbol.ops = fake_ops
    pk_c, pk_d = fake_ops.pack_conv_weight, fake_ops.pack_deconv_weight
    fake_ops.pack_conv_weight = lambda w, cin=None, cout=None, dtype=None: pk_c(w, cin, cout, torch.float64)
    fake_ops.pack_deconv_weight = lambda w, cin=None, cout=None, dtype=None: pk_d(w, cin, cout, torch.float64)

    def get(self, key, shape, device):
        t = self.bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = self.bufs[key] = torch.zeros(tuple(shape), dtype=torch.float64)
        return t
    dla._BufferPool.get = get
    to_pad = fake_ops.to_nhwc_padded

    def to_nhwc_padded(x, channels=None):      # the real helper rounds to bf16: keep float64 here
        N, C, H, W = x.shape
        out = torch.zeros((N, H + 2, W + 2, channels or C), dtype=torch.float64)
        out[:, 1:H + 1, 1:W + 1, :C] = x.permute(0, 2, 3, 1)
        return out
    fake_ops.to_nhwc_padded = to_nhwc_padded
    fake_ops.from_nhwc_padded = lambda y, channels=None: (y[:, 1:-1, 1:-1, :channels] if channels else y[:, 1:-1, 1:-1]).permute(0, 3, 1, 2).contiguous()

    tmp = pathlib.Path(tempfile.mkdtemp())
    with trc.drop_in():
        pcx = types.ModuleType("processing_cxx")
        pcx.assign3D_v2 = lambda pc, bbox, ctr, rad, mask, nlz, *f: orc.assign3d_v2(pc, bbox, ctr, rad, mask, nlz, *f).reshape(-1, 1)
        pcx.get_point_num = lambda inds: orc.get_point_num(inds).reshape(-1, 1)
        sys.modules["processing_cxx"] = pcx
        cfg = importlib.import_module("config.rangedet.rangedet_veh_wo_aug_4_18e")
        out = cfg.get_config(is_train=False)
        pModel, transform, data_name = out[6], out[9], out[10]
        rec = trc._raw_record(tmp)
        for t in transform:
            t.apply(rec)
        batch = {k: torch.from_numpy(np.ascontiguousarray(rec[k], dtype=np.float32)[None]) for k in data_name}
        test_sym = pModel.test_symbol
    g = torch.Generator().manual_seed(3)
    P = dla_ref.make_params(seed=0, device="cpu")
    for k in P:
        if k.endswith("_gamma"):
            P[k] = 1 + 0.2 * torch.randn(P[k].shape, generator=g)
    # a "trained-looking" state: moving statistics := the batch statistics of this frame (one training-mode pass of the
    # reference graph with BatchNorm momentum 0), so that the inference graph sees normalised activations
    warm = ref_graph.backbone_head(P, batch["input_data"], batch["coord_s1"], training=True, bn_momentum=0.0)
    P.update(warm["moving"])
    P64 = {k: v.double() for k, v in P.items()}
    data, coord = batch["input_data"].double(), batch["coord_s1"].double()
    # (b) ours
    ex = symbol._TestExecutor.__new__(symbol._TestExecutor)
    ex.sym, ex.pre_n, ex.post_n, ex.nms_thr, ex.wnms = test_sym, 2000, 200, 0.2, True
    cls, reg = dla.RangeRpnHead(P64, "cpu").get_fpn_output(dla.DLABackbone(P64, "cpu").get_rpn_feature(data, coord))
    score, boxes, _ = ex.predict(cls, reg, batch)
    # (a) the reference graph
    r = ref_graph.backbone_head(P64, data, coord, training=False)
    sc_r, box_r = ref_graph.fpn_prediction([c.float() for c in r["cls"]], [d.float() for d in r["reg"]],
                                           [batch["pc_vehicle_frame_s%d" % s] for s in (1, 2, 4)],
                                           [batch["range_image_mask_s%d" % s].reshape(1, -1) for s in (1, 2, 4)], 2000)
    rms = lambda a, b: float(((a.double() - b.double()) ** 2).mean().sqrt() / (b.double() ** 2).mean().sqrt())
    print(json.dumps({"head_rms": [rms(a, b) for a, b in zip(cls + reg, r["cls"] + r["reg"])],
                      "score_max_abs_diff": float((score - sc_r).abs().max()), "top_score": float(sc_r.max()),
                      "ranks_with_different_box": int(((boxes - box_r).abs().amax(-1) > 1e-3).sum()),
                      "boxes_as_sets_max_abs_diff": float(torch.cdist(boxes[0].double(), box_r[0].double(), p=float("inf")).min(1).values.max()),
                      "boxes_finite": bool(torch.isfinite(box_r).all()),
                      "n": int(score.numel()), "distinct_scores": int(torch.unique(sc_r).numel())}))


if __name__ == "__main__":
    main()
