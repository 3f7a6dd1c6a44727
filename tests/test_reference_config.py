"""the reference's UNMODIFIED config file builds its graphs through this package's classes.

config/rangedet/rangedet_veh_wo_aug_4_18e.py is imported verbatim from /root/reference with
  rangedet.symbol.head.builder       -> rangedet_b200.symbol (RangeRCNN, RangeRpnHead)
  rangedet.symbol.backbone.dla_backbone -> rangedet_b200.symbol (DLABackbone)
  processing_cxx                     -> rangedet_b200.processing_cxx
by the shim this package SHIPS (rangedet_b200/shim: install() / drop_in()), and the rest of the reference
(rangedet.core.input, rangedet.core.detection_metric, utils) as is.  MXNet is absent here, so the shim also supplies the
three host-side base classes those files subclass (shim/mxnet_host.py); nothing under oracle/ is imported.  `get_config()` must return
our graph objects, and everything the config says about them -- metric output names, data / label names, the
optimizer block -- must line up with what they expose.  Needs /root/reference: skipped elsewhere."""
import contextlib
import importlib
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "config")), reason="/root/reference not present")


@contextlib.contextmanager
def drop_in(processing_cxx=None):
    """The SHIPPED shim (rangedet_b200/shim): no test-side patching, nothing from oracle/."""
    from rangedet_b200 import shim
    with shim.drop_in(REF, processing_cxx=processing_cxx):
        yield


@pytest.mark.parametrize("name", ["rangedet_veh_wo_aug_4_18e", "rangedet_veh_wo_aug_all_36e", "rangedet_ped_wo_aug_4_18e"])
def test_unmodified_config_builds_our_graphs(name):
    from rangedet_b200 import symbol
    with drop_in():
        cfg = importlib.import_module("config.rangedet." + name)
        (pGen, pKv, pRpn, pRoi, pBbox, pDataset, pModel, pOpt, pTest, transform, data_name, label_name, metric_list,
         pLabelMap) = cfg.get_config(is_train=True)
        train_sym = pModel.train_symbol
        assert isinstance(train_sym, symbol.TrainSymbol) and pModel.test_symbol is None
        # the metrics the config registers read exactly the outputs our train symbol lists (tools/train.py binds them by name)
        assert sorted(o for m in metric_list for o in m.output_names) == sorted(train_sym.list_outputs())
        # every array the loader is told to provide is a graph input of ours
        inputs = train_sym.list_inputs()
        assert set(data_name) | set(label_name) <= set(inputs), sorted((set(data_name) | set(label_name)) - set(inputs))
        sh = train_sym.infer_shape()
        assert sh["input_data"] == (pGen.batch_image, 8, 64, 2656) and all(n in sh for n in data_name + label_name)
        # loss hyper-parameters and optimiser block as our bind() consumes them
        hyp = train_sym.head.loss_hyper()
        assert hyp["scale_loss_shift"] == 128.0 and hyp["smooth_l1_scalar"] == 3.0 and hyp["iou_type"] == "bev"
        assert pOpt.optimizer.type == "sgd" and pOpt.optimizer.clip_gradient == 35 and pOpt.optimizer.momentum == 0.9
        names = [type(t).__name__ for t in transform]                      # the loader pipeline is the reference's own
        assert names.index("Bbox3dAssigner") < names.index("GenerateTarget") < names.index("GenerateFPNTarget")
        # test-time graph from the same file
        cfg_t = cfg.get_config(is_train=False)
        test_sym = cfg_t[6].test_symbol
        assert isinstance(test_sym, symbol.TestSymbol) and set(cfg_t[10]) <= set(test_sym.list_inputs())
        assert cfg_t[8].nms.wnms is True and cfg_t[8].nms.thr_lo == 0.1


def test_drop_in_context_restores_the_interpreter():
    before = set(sys.modules)
    with drop_in():
        importlib.import_module("mxnext.complicate")
        assert "mxnext" in sys.modules
    assert not [k for k in sys.modules if k.split(".")[0] in ("mxnext", "mxnet", "config", "utils")]
    assert "processing_cxx" not in sys.modules


def test_shim_is_product_code_and_needs_no_oracle():
    """The drop-in lives in the package and resolves every module path the config imports without oracle/."""
    import subprocess
    code = ("import sys, importlib; import rangedet_b200.shim as shim; shim.install(%r); "
            "cfg = importlib.import_module('config.rangedet.rangedet_veh_wo_aug_4_18e'); out = cfg.get_config(True); "
            "from rangedet_b200 import symbol; assert isinstance(out[6].train_symbol, symbol.TrainSymbol); "
            "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')], 'oracle imported'; "
            "import rangedet.symbol.backbone.meta_kernel as mk; from rangedet_b200.meta_kernel import MetaKernel; "
            "assert mk.MetaKernel is MetaKernel; print('ok')" % REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_shim_refuses_a_normalizer_the_kernels_do_not_implement():
    from rangedet_b200 import symbol
    from rangedet_b200.shim.mxnext_complicate import normalizer_factory
    with drop_in():
        cfg = importlib.import_module("config.rangedet.rangedet_veh_wo_aug_4_18e")
        pBackbone_cls = None
        out = cfg.get_config(is_train=True)
        backbone = out[6].train_symbol.backbone
        p = backbone.p
        assert p.normalizer.type == "localbn"

        class P2(p):
            normalizer = normalizer_factory(type="syncbn", ndev=8)

        with pytest.raises(NotImplementedError):
            symbol.DLABackbone(P2)
    with pytest.raises(KeyError):
        normalizer_factory(type="nonsense")


def _raw_record(tmp_path, seed=0):
    """A synthetic raw roidb record + the npz LoadRecord reads (rangedet/core/input.py:24-42)."""
    import numpy as np
    from rangedet_b200 import synth
    pc, mask, b7, c24 = synth.assign_frame(n_vehicles=20, seed=seed)
    H, W = 64, 2650
    rng = np.random.default_rng(seed)
    pc = pc.reshape(H, W, 3)
    rim = np.zeros((H, W, 4), np.float32)
    rng_val = np.linalg.norm(pc, axis=2)
    valid = mask.reshape(H, W) > 0
    rim[..., 0] = np.where(valid, rng_val, -1.0)
    rim[..., 1] = rng.uniform(0, 1, (H, W)) * valid
    rim[..., 2] = rng.uniform(0, 1, (H, W)) * valid
    rim[..., 3] = -1.0
    url = str(tmp_path / "frame.npz")
    np.savez(url, pc_vehicle_frame=pc, range_image=rim, inclination=np.linspace(-0.31, 0.04, H).astype(np.float32),
             azimuth=np.linspace(np.pi, -np.pi, W).astype(np.float32))
    M = b7.shape[0]
    return {"pc_url": url, "gt_class": np.ones((M,), np.int64), "gt_bbox_yaw": b7[:, 6].copy(), "gt_bbox_csa": b7.copy(),
            "gt_bbox_imu": c24.reshape(M, 8, 3).copy(), "meta_data": np.zeros((M, 4)), "points_in_box": np.full((M,), 10.0),
            "rec_id": np.array([0])}


def test_reference_loader_pipeline_output_fits_our_graph_inputs(tmp_path):
    """The config's own transform list (LoadRecord ... TransAndReshape, rangedet/core/input.py) executed on a synthetic
    raw frame: the record it produces must match, name by name and shape by shape, what TrainSymbol.infer_shape()
    declares and GraphedTrainStep.set_targets() copies.  (processing_cxx = the CPU restatement here: the product
    module has no CPU path.)"""
    import numpy as np
    from oracle import oracle
    orc = oracle()
    pcx = types.ModuleType("processing_cxx")
    pcx.assign3D_v2 = lambda pc, bbox, ctr, rad, mask, nlz, *f: orc.assign3d_v2(pc, bbox, ctr, rad, mask, nlz, *f).reshape(-1, 1)
    pcx.get_point_num = lambda inds: orc.get_point_num(inds).reshape(-1, 1)
    with drop_in(processing_cxx=pcx):
        cfg = importlib.import_module("config.rangedet.rangedet_veh_wo_aug_4_18e")
        out = cfg.get_config(is_train=True)
        pModel, transform, data_name, label_name = out[6], out[9], out[10], out[11]
        rec = _raw_record(tmp_path)
        for t in transform:
            t.apply(rec)
        sh = pModel.train_symbol.infer_shape(batch_image=1)
        for n in data_name + label_name:
            assert n in rec, n
            assert tuple(rec[n].shape) == tuple(sh[n][1:]), (n, rec[n].shape, sh[n])
            assert np.asarray(rec[n]).dtype in (np.float32, np.float64), (n, rec[n].dtype)
        # content sanity of the record the kernels will see
        assert rec["gt_bbox_veh_for_iou_pred"].shape == (200, 8) and np.allclose(rec["gt_bbox_veh_for_iou_pred"][-1], [0, 0, 0, 1e-3, 1e-3, 1e-3, 1e-3, 0])
        assert (rec["rpn_reg_weight_s1"] > 0).any() and not rec["rpn_reg_target_s1"][:, :, 2650:].any()
        assert set(np.unique(rec["range_image_mask_s1"])) <= {0.0, 1.0}
