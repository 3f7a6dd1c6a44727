"""Whole drop-in chain on the CPU, in THIS container (needs /root/reference; ~2 min):

  the reference's UNMODIFIED config file  ->  its own loader transform list on a synthetic raw frame  ->
    (a) the reference's own train graph (dla_backbone.py + head/builder.py + loss.py) executed eagerly through
        oracle/mx_eager.py, fp32
    (b) this package: symbol.RangeRCNN(...).get_train_symbol(...).bind(capture=False) over the plain-torch emulation
        of the kernel API (tests/fake_ops.py; bf16 storage like the kernels)
  on the SAME parameters and the SAME record: per-level loss sums and head outputs side by side.

    python scripts/dropin_chain_cpu.py            # emulated kernels with bf16 storage, like the GPU path
    python scripts/dropin_chain_cpu.py --exact    # float64 activations / operands: isolates the LOGIC from the storage format

TEST INFRASTRUCTURE (imports oracle/ and tests/); prints a JSON line.
"""
import importlib
import json
import os
import sys
import tempfile
import pathlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fake_ops  # noqa: E402
import test_reference_config as trc  # noqa: E402
from oracle import dla_ref, oracle, ref_graph  # noqa: E402
from rangedet_b200 import train  # noqa: E402


def main():
    import types
    torch.manual_seed(0)
    orc = oracle()
    tmp = pathlib.Path(tempfile.mkdtemp())
    with trc.drop_in():
        pcx = types.ModuleType("processing_cxx")
        pcx.assign3D_v2 = lambda pc, bbox, ctr, rad, mask, nlz, *f: orc.assign3d_v2(pc, bbox, ctr, rad, mask, nlz, *f).reshape(-1, 1)
        pcx.get_point_num = lambda inds: orc.get_point_num(inds).reshape(-1, 1)
        sys.modules["processing_cxx"] = pcx
        cfg = importlib.import_module("config.rangedet.rangedet_veh_wo_aug_4_18e")
        out = cfg.get_config(is_train=True)
        pModel, pOpt, transform, data_name, label_name = out[6], out[7], out[9], out[10], out[11]
        rec = trc._raw_record(tmp)
        for t in transform:
            t.apply(rec)
        batch = {k: np.ascontiguousarray(rec[k], dtype=np.float32)[None] for k in data_name + label_name}
        train_sym = pModel.train_symbol
        P = dla_ref.make_params(seed=0, device="cpu")
        train.ops = fake_ops
        step = train_sym.bind({k: v.clone() for k, v in P.items()}, batch_image=1, optimizer=pOpt.optimizer, device="cpu")
    return P, batch, step


if __name__ == "__main__":
    exact = "--exact" in sys.argv
    if exact:
        fake_ops.set_exact(True)
        _pack = train.pack_operand
        train.pack_operand = lambda w, kind, ci_p, co_p, S=1, dtype=None: _pack(w, kind, ci_p, co_p, S, torch.float64)

        def _get(self, key, shape):
            t = self.bufs.get(key)
            if t is None or tuple(t.shape) != tuple(shape):
                t = self.bufs[key] = torch.zeros(tuple(shape), dtype=torch.float64)
            return t
        train._Pool.get = _get
    # bind() builds a captured step by default; this script needs the eager mode
    orig = train.GraphedTrainStep

    def eager(*a, **k):
        k.update(capture=False, overlap_wgrad=False)
        return orig(*a, **k)
    train.GraphedTrainStep = eager
    P, batch, step = main()
    step.set_targets(batch)
    data, coord = torch.from_numpy(batch["input_data"]), torch.from_numpy(batch["coord_s1"])
    step.set_lr(0.0)
    ours = step.train_step(data, coord)
    if exact:
        r = ref_graph.backbone_head({k: v.double() for k, v in P.items()}, data.double(), coord.double(), training=True, targets=batch)
    else:
        r = ref_graph.backbone_head(P, data, coord, training=True, targets={k: v for k, v in batch.items()})
    res = {"mode": "float64 (logic only)" if exact else "bf16 storage (like the kernels)", "ours_cls": [float(o["cls_loss"].sum()) for o in ours], "ref_cls": [float(x.sum()) for x in r["cls_loss"]],
           "ours_reg": [float(o["reg_loss"].sum()) for o in ours], "ref_reg": [float(x.sum()) for x in r["reg_loss"]],
           "head_rms": [float(((a.double() - b.double()) ** 2).mean().sqrt() / (b.double() ** 2).mean().sqrt())
                        for a, b in zip(step.out[0] + step.out[1], r["cls"] + r["reg"])]}
    print(json.dumps(res))
